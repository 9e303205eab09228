"""LayerNorm forward / backward kernels alone at the encoder shape (rows = 32000, d = 512, fp32 in / out), CUDA-event timed,
L2 flushed between launches; per option value.  python tools/bench_ln.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
lib = stb._lib.load()
dev = torch.device("cuda", 0)
rows, d = 32000, 512
p = lambda t: None if t is None else t.data_ptr()
S = lambda: torch.cuda.current_stream().cuda_stream
a = torch.randn(rows, d, device=dev); g = torch.ones(d, device=dev); b = torch.zeros(d, device=dev)
out = torch.empty_like(a); mean = torch.empty(rows, device=dev); rstd = torch.empty(rows, device=dev)
dy = torch.randn(rows, d, device=dev); dz = torch.empty_like(a); dg = torch.zeros(d, device=dev); db = torch.zeros(d, device=dev); ds = torch.zeros(d, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize(); tot = 0.0
    for _ in range(reps):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); tot += e0.elapsed_time(e1)
    return tot / reps * 1e3
fwd = lambda: stb._lib.check(lib.st_add_ln_fwd(p(a), None, p(g), p(b), p(out), None, p(mean), p(rstd), rows, d, 1e-6, 0, 0.0, 0, S()))
bwd = lambda pd=0.0: stb._lib.check(lib.st_add_ln_bwd(p(dy), p(a), p(mean), p(rstd), p(g), p(dz), p(dg), p(db), p(ds), rows, d, 0, pd, 7, S()))
for opt in ("ln_fwd_registers", "ln_bwd_registers"):
    for v in (1, 0):
        lib.st_set_option(opt.encode(), v)
        if "fwd" in opt:
            us = timeit(fwd); gb = rows * d * 8 / 1e9
        else:
            us = timeit(bwd); gb = rows * d * 12 / 1e9
        print(f"{opt}={v}: {us:7.1f} us  {gb / (us * 1e-6) / 1e3:6.2f} TB/s")
    lib.st_set_option(opt.encode(), 0)
us = timeit(lambda: bwd(0.1)); print(f"bwd with dropout 0.1: {us:7.1f} us")
