"""CPU ORACLE (test infrastructure only) — beam decoding as the reference describes it, on the CPU oracle model.

Restates transformer/Decode.py:48-179 (decode_batch: encode once, repeat per beam, re-run the decoder on the FULL
prefix every step, feed the last position's log-probabilities to the beams) and transformer/Beam.py:43-74
(Beam.advance: add running scores, top-k over beam x vocab, back-pointer = index // vocab, symbol = index % vocab,
done when the best hypothesis ends in EOS) on top of oracle/model_port.forward.  The reference files themselves are
stale (undefined `prob_projection`, old Transformer constructor, float division for the back-pointer — SURVEY.md
§2/§8f); this is their evident intent with log_softmax as the probability projection.  O(L^2), CPU, any float dtype.
"""
import torch

from . import model_port

PAD, UNK, BOS, EOS = 0, 1, 2, 3   # transformer/Constants.py:1-4


def step_logits(P: dict, cfg: dict, inputs, in_len, prefix):
    """Logits for the next symbol after each row of `prefix` (N, t+1); inputs/in_len already repeated per row."""
    n, t1 = prefix.shape
    tgt_len = torch.full((n,), t1, dtype=torch.int64)
    return model_port.forward(P, cfg, inputs, in_len, prefix, tgt_len)[:, -1, :]


def beam_advance(scores, word_lk, first: bool):
    """One Beam.advance (Beam.py:43-74) for a batch of utterances.  scores (B, beam) running scores, word_lk (B, beam, V)
    log-probabilities of the next symbol.  Returns (new scores, prev_k, next_y), each (B, beam), best first: at the first
    position only hypothesis 0 is expanded (Beam.py:49-52, all hypotheses are identical); afterwards the `beam` best of the
    flattened beam x word array (Beam.py:56-59), prev_k = id // num_words (Beam.py:66, integer division as intended),
    next_y = id - prev_k * num_words (Beam.py:68).  PINNED: tests/golden/beam_advance.npz holds what the reference's own
    Beam class produces for seeded inputs (oracle/make_golden_beam.py)."""
    B, beam, V = word_lk.shape
    cand = word_lk[:, :1] if first else word_lk + scores.unsqueeze(2)           # Beam.py:49-52
    best, idx = cand.reshape(B, -1).topk(beam, dim=1)                           # Beam.py:56-59
    prev_k = idx // V                                                           # Beam.py:66
    return best, prev_k, idx - prev_k * V                                       # Beam.py:68


def hypothesis(prev_ks, next_ys, k: int):
    """Beam.get_hypothesis (Beam.py:99-116) for one utterance: walk the back-pointers from beam position k.
    prev_ks / next_ys: per-step (beam,) tensors (next_ys WITHOUT the initial BOS row)."""
    hyp = []
    for j in range(len(prev_ks) - 1, -1, -1):
        hyp.append(int(next_ys[j][k]))
        k = int(prev_ks[j][k])
    return hyp[::-1]


def beam_search(P: dict, cfg: dict, inputs, in_len, beam: int, max_len: int, n_best: int = 1, eos: int = EOS, bos: int = BOS):
    B, V = inputs.size(0), cfg["vocab_size"]
    rep_in = inputs.repeat_interleave(beam, 0)           # Decode.py:62-68
    rep_len = in_len.repeat_interleave(beam, 0)
    prefix = torch.full((B * beam, 1), int(bos), dtype=torch.int64)
    scores = torch.zeros(B, beam, dtype=P["tgt_word_proj.weight"].dtype)
    done = torch.zeros(B, dtype=torch.bool)
    for t in range(max_len):
        logp = torch.log_softmax(step_logits(P, cfg, rep_in, rep_len, prefix), -1).view(B, beam, V)
        best, prev_k, y = beam_advance(scores, logp, first=(t == 0))         # Beam.py:43-74
        keep = done.unsqueeze(1)
        prev_k = torch.where(keep, torch.arange(beam).expand(B, -1), prev_k)
        y = torch.where(keep, torch.full_like(y, PAD), y)
        scores = torch.where(keep, scores, best)
        parent = (torch.arange(B).unsqueeze(1) * beam + prev_k).reshape(-1)
        prefix = torch.cat([prefix[parent], y.reshape(-1, 1)], 1)            # Beam.get_tentative_hypothesis
        done = done | (y[:, 0] == eos)                                       # Beam.py:70-72
        if bool(done.all()):
            break
    order = scores.sort(dim=1, descending=True)
    hyps = []
    for b in range(B):
        per = []
        for k in order.indices[b, :n_best].tolist():
            per.append([int(x) for x in prefix[b * beam + k, 1:].tolist() if x != PAD])
        hyps.append(per)
    return hyps, order.values[:, :n_best]
