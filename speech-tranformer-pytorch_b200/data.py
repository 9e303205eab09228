"""Synthetic filterbank batches in the layout of the reference's Dataset.__getitem__ (Dataset.py:34-51,
56-68): inputs (B, T_max, F) fp32 ~ N(0,1) (post-CMVN statistics, Dataset.py:89-92) with zero-filled tails,
targets = [BOS] + labels, ground_truth = labels + [BOS] (sic, Dataset.py:36-37), PAD = 0 padding, int64
length vectors.  There is no network for datasets, so this is what bench.py and the smoke test feed."""
import torch

PAD, UNK, BOS, EOS = 0, 1, 2, 3


def synthetic_batch(batch: int, t_max: int, l_max: int, feat: int, vocab: int, seed: int = 2018, fixed_len: bool = True,
                    t_min: int = 1, l_min: int = 10, pin: bool = False):
    g = torch.Generator().manual_seed(seed)
    if fixed_len:
        in_len = torch.full((batch,), t_max, dtype=torch.int64)
    else:
        in_len = torch.randint(max(t_min, 1), t_max + 1, (batch,), generator=g)
        in_len[0] = t_max
    tgt_len = torch.randint(min(l_min, l_max), l_max + 1, (batch,), generator=g)
    tgt_len[0] = l_max
    inputs = torch.randn(batch, t_max, feat, generator=g)
    targets = torch.zeros(batch, l_max, dtype=torch.int64)
    truth = torch.zeros(batch, l_max, dtype=torch.int64)
    for b in range(batch):
        inputs[b, int(in_len[b]):] = 0
        n = int(tgt_len[b]) - 1
        labels = torch.randint(4, vocab, (n,), generator=g)
        targets[b, 0] = BOS
        targets[b, 1:n + 1] = labels
        truth[b, :n] = labels
        truth[b, n] = BOS
    out = (inputs, targets, in_len, tgt_len, truth)
    if pin and torch.cuda.is_available():
        out = tuple(t.pin_memory() for t in out)
    return out


class DevicePrefetcher:
    """Double-buffered host -> device input pipeline (what a DataLoader with pin_memory feeds, train.py:29-36): `submit`
    starts the copy of a pinned host batch on a side stream, `get` hands the tensors to the compute stream once the
    copy has landed.  Submitting batch k+1 before running step k hides the PCIe transfer under the step."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self._pending = None

    def submit(self, host_batch) -> None:
        if self._pending is not None:
            raise RuntimeError("DevicePrefetcher: the previous batch has not been taken")
        with torch.cuda.stream(self.stream):
            batch = [t.to(self.device, non_blocking=True) for t in host_batch]
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self._pending = (batch, ev)

    def get(self):
        if self._pending is None:
            raise RuntimeError("DevicePrefetcher: nothing submitted")
        batch, ev = self._pending
        self._pending = None
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(ev)
        for t in batch:
            t.record_stream(cur)       # allocated on the side stream, consumed on the compute stream
        return batch
