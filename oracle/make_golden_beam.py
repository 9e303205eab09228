"""Golden vectors for the beam-search bookkeeping, produced by the REFERENCE's own Beam class (transformer/Beam.py:43-74,
from oracle/_ref with the one-line floor-division patch of oracle/make_ref.py).

    python oracle/make_ref.py && python oracle/make_golden_beam.py      # build container only (needs /root/reference)

Writes tests/golden/beam_advance.npz: for several (beam, vocabulary, seed) cases a sequence of seeded log-probability
tables is fed to Beam.advance step by step; recorded per step: the scores, back-pointers and symbols the reference keeps,
its `done` flag, and at the end every hypothesis from Beam.get_hypothesis.  oracle/decode_port.beam_advance / hypothesis
are held to these by tests/test_oracle_golden.py, which pins the oracle the GPU beam search is compared with."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "_ref"))
from transformer.Beam import Beam          # noqa: E402  the reference's class
import transformer.Constants as Constants  # noqa: E402

CASES = [(4, 31, 6, 1), (10, 57, 8, 2), (1, 17, 5, 3), (10, 1201, 4, 4), (3, 9, 7, 5)]   # beam, vocab, steps, seed


def main():
    out = {"cases": np.array(CASES), "constants": np.array([Constants.PAD, Constants.UNK, Constants.BOS, Constants.EOS])}
    for ci, (beam, V, steps, seed) in enumerate(CASES):
        g = torch.Generator().manual_seed(seed)
        b = Beam(beam, torch.device("cpu"))
        lk, scores, prev, ys, done = [], [], [], [], []
        for t in range(steps):
            word_lk = torch.log_softmax(3.0 * torch.randn(beam, V, generator=g), dim=-1)
            if t == steps - 2 and ci % 2 == 0:      # let some cases finish early: make EOS the best continuation of the top beam
                word_lk[int(b.scores.argmax()) if t > 0 else 0, Constants.EOS] = 0.0
            lk.append(word_lk.numpy().copy())
            finished = b.advance(word_lk)
            scores.append(b.scores.numpy().copy())
            prev.append(b.prev_ks[-1].numpy().copy())
            ys.append(b.next_ys[-1].numpy().copy())
            done.append(bool(finished))
            if finished:
                break
        hyps = np.array([[int(x) for x in b.get_hypothesis(k)] for k in range(beam)])
        out[f"c{ci}.word_lk"] = np.stack(lk)
        out[f"c{ci}.scores"] = np.stack(scores)
        out[f"c{ci}.prev_k"] = np.stack(prev)
        out[f"c{ci}.next_y"] = np.stack(ys)
        out[f"c{ci}.done"] = np.array(done)
        out[f"c{ci}.hyps"] = hyps
    path = os.path.join(os.path.dirname(HERE), "tests", "golden", "beam_advance.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, {k: v.shape for k, v in out.items() if k.startswith("c0")})


if __name__ == "__main__":
    main()
