"""Timeline of CTA (0,0,0) of the attention dK/dV kernel at the encoder shape (option attn_trace):
cycles relative to the first event, per 64-query tile.  python tools/trace_attn.py [res_smem]"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
F = stb.functional
lib = stb._lib.load()
dev = "cuda:0"
torch.manual_seed(0)
B, H, L, dk = 32, 8, 1000, 64
d = H * dk
q, k, v = (torch.randn(B, L, d, device=dev).requires_grad_() for _ in range(3))
go = torch.randn(B, L, d, device=dev)
if len(sys.argv) > 1:
    lib.st_set_option(b"attn_dkv_res_smem", 1)
out, _ = F.attention_core(q, k, v, None, n_head=H, dropout_p=0.1, seed=7)
out.backward(go, retain_graph=True)
lib.st_set_option(b"attn_trace", 1)
out.backward(go, retain_graph=True)
torch.cuda.synchronize()
lib.st_set_option(b"attn_trace", 0)
TT, EV = 24, 6
buf = (C.c_uint64 * (2 * TT * EV))()
n = lib.st_debug_read_trace(buf, 2 * TT * EV)
t0 = min(x for x in buf if x)
print("MMA warp : tile | A: wait ld_full start, ld_full ok, A issued+commit | B: wait ds_full start, ds_full ok, B issued+commit")
for t in range(16):
    r = [buf[(0 * TT + t) * EV + e] - t0 for e in range(EV)]
    print(f"  {t:2d} | {r[0]:7d} {r[1]:7d} {r[2]:7d} | {r[3]:7d} {r[4]:7d} {r[5]:7d}")
print("compute thread 0: tile | wait s_full start, s_full ok, tmem ld done, math done, tmem st done, arrived")
prev = None
for t in range(16):
    r = [buf[(1 * TT + t) * EV + e] - t0 for e in range(EV)]
    print(f"  {t:2d} | {r[0]:7d} {r[1]:7d} {r[2]:7d} {r[3]:7d} {r[4]:7d} {r[5]:7d}   wait {r[1]-r[0]:5d} ld {r[2]-r[1]:5d} math {r[3]-r[2]:5d} st {r[4]-r[3]:5d}")
# ---- per-CTA wall clock: duration of every CTA and idle gaps per SM
NC = 2048
big = (C.c_uint64 * (2 * TT * EV + NC * 4))()
lib.st_debug_read_trace(big, 2 * TT * EV + NC * 4)
recs = [(big[2 * TT * EV + 4 * i], big[2 * TT * EV + 4 * i + 1], big[2 * TT * EV + 4 * i + 2]) for i in range(NC)]
recs = [r for r in recs if r[0] and r[1]]
g0 = min(r[0] for r in recs)
durs = sorted(r[1] - r[0] for r in recs)
print(f"CTAs {len(recs)}  duration ns: min {durs[0]} median {durs[len(durs)//2]} p90 {durs[int(.9*len(durs))]} max {durs[-1]}  kernel span {max(r[1] for r in recs) - g0} ns")
by = {}
for s_, e_, sm in recs: by.setdefault(sm, []).append((s_ - g0, e_ - g0))
sm0 = sorted(by)[0]
seq = sorted(by[sm0])
print(f"SM {sm0}: {len(seq)} CTAs:", [(a, b - a) for a, b in seq][:16])
gaps = [seq[i + 1][0] - seq[i][1] for i in range(len(seq) - 1)]
print("gaps between consecutive CTAs on that SM (ns):", gaps)
