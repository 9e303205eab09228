"""Build oracle/_ref/: a scratch copy of the reference's `transformer/` package with the minimal
compatibility patch that lets the WHOLE model run on a current PyTorch.

    python oracle/make_ref.py            # needs /root/reference (build container only)

oracle/_ref/ is git-ignored build output (it is NOT part of this repository's sources); it travels to
the GPU box with the working tree so that `bench.py --impl reference` and the CPU-baseline leg can time
the reference's own code on the box's host cores.  Every patch is an exact one-line substitution that
must match exactly once; nothing else is touched.

  Attention.py:94   `output + v` -> `output + q`       cross-attention residual (len_q != len_k cannot add v)
  Models.py:87      position_enc(dec_input) -> dec_input + position_enc(outputs_pos)
  Models.py:90,92,97  masks built from lengths (outputs_pos), not token ids (outputs_data)
  Models.py:102     the layer returns (out, (slf, enc)) — unpack what Layers.py:44 actually returns
  Utils.py:52,66    uint8 masks -> bool (masked_fill_ rejects uint8 on torch >= 2)
  Beam.py:66        `best_scores_id / num_words` -> `//` (true division makes the back-pointers floats on torch >= 1.5)
"""
import os
import shutil
import sys

REF = os.environ.get("ST_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")

PATCHES = {
    "Attention.py": [
        ("norm_output = self.layernorm(output + v)", "norm_output = self.layernorm(output + q)"),
    ],
    "Models.py": [
        ("dec_input = self.position_enc(dec_input)", "dec_input = dec_input + self.position_enc(outputs_pos)"),
        ("dec_slf_attn_pad_mask = padding_info_mask(\n            outputs_data, outputs_data)",
         "dec_slf_attn_pad_mask = padding_info_mask(\n            outputs_pos, outputs_pos)"),
        ("dec_slf_attn_sub_mask = feature_info_mask(outputs_data)", "dec_slf_attn_sub_mask = feature_info_mask(outputs_pos)"),
        ("dec_enc_attn_pad_mask = padding_info_mask(\n            outputs_data, input_pos)",
         "dec_enc_attn_pad_mask = padding_info_mask(\n            outputs_pos, input_pos)"),
        ("dec_output, dec_slf_attn, dec_enc_attn = dec_layer(", "dec_output, (dec_slf_attn, dec_enc_attn) = dec_layer("),
    ],
    "Beam.py": [
        ("prev_k = best_scores_id / num_words", "prev_k = best_scores_id // num_words"),
    ],
    "Utils.py": [
        ("pad_attn_mask = torch.from_numpy(mask_mat).unsqueeze(1)", "pad_attn_mask = torch.from_numpy(mask_mat).bool().unsqueeze(1)"),
        ("subsequent_mask = torch.from_numpy(subsequent_mask)", "subsequent_mask = torch.from_numpy(subsequent_mask).bool()"),
        ("import editdistance\n", ""),
        ("import matplotlib.pyplot as plt\n", ""),
    ],
}


def main() -> int:
    src = os.path.join(REF, "transformer")
    if not os.path.isdir(src):
        print(f"make_ref: {src} not found — nothing built (expected on the GPU box)")
        return 0
    pkg = os.path.join(DST, "transformer")
    shutil.rmtree(DST, ignore_errors=True)
    os.makedirs(pkg)
    for name in sorted(os.listdir(src)):
        if not name.endswith(".py"):
            continue
        with open(os.path.join(src, name), encoding="utf-8") as f:
            text = f.read()
        for old, new in PATCHES.get(name, []):
            if text.count(old) != 1:
                raise SystemExit(f"make_ref: pattern must match exactly once in {name}: {old!r} (found {text.count(old)})")
            text = text.replace(old, new)
        with open(os.path.join(pkg, name), "w", encoding="utf-8") as f:
            f.write(text)
    with open(os.path.join(DST, "README"), "w") as f:
        f.write("Scratch copy of the reference's transformer/ package + compat patch (oracle/make_ref.py). Not source.\n")
    print(f"make_ref: wrote {pkg}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
