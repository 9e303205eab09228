"""Timeline of thread 0 of CTA (0,0,0) of the attention forward kernel at the encoder shape (option attn_trace).
python tools/trace_attn_fwd.py"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
F = stb.functional; lib = stb._lib.load(); dev = "cuda:0"
torch.manual_seed(0)
B, H, L, dk = 32, 8, 1000, 64
q, k, v = (torch.randn(B, L, H * dk, device=dev) for _ in range(3))
with torch.no_grad():
    F.attention_core(q, k, v, None, n_head=H, dropout_p=0.1, seed=7)
    lib.st_set_option(b"attn_trace", 1)
    F.attention_core(q, k, v, None, n_head=H, dropout_p=0.1, seed=7)
    torch.cuda.synchronize(); lib.st_set_option(b"attn_trace", 0)
TT, EV = 16, 8
buf = (C.c_uint64 * (TT * EV))(); lib.st_debug_read_fwd_trace(buf, TT * EV)
t0 = min(x for x in buf if x)
print("tile | wait S start, S ok, ld done, max+sync done, probs+st done, sync2 done, V ok, issued  || S wait, ld, max, probs, sync2, issue")
for t in range(8):
    r = [buf[t * EV + e] - t0 for e in range(EV)]
    print(f" {t:2d} | " + " ".join(f"{x:7d}" for x in r) + f" || {r[1]-r[0]:5d} {r[2]-r[1]:5d} {r[3]-r[2]:5d} {r[4]-r[3]:5d} {r[5]-r[4]:5d} {r[7]-r[5]:5d}")
