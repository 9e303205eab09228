"""Attention core kernels at the encoder shape (B=32, h=8, L=1000, d_k=64) and the cross-attention shape (Lq=50), fp16:
per-class CUDA-event times from the library's profile (side streams off so intervals do not overlap).
python tools/bench_attn16.py [fp16|bf16] [dropout]"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
F = stb.functional; lib = stb._lib.load()
dt = torch.bfloat16 if (len(sys.argv) > 1 and sys.argv[1] == "bf16") else torch.float16
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
dev = "cuda:0"; H, dk = 8, 64
lib.st_set_option(b"side_streams", 0)
names = {1: "attn_fwd", 2: "attn_bwd_dkv", 3: "attn_bwd_dq", 4: "attn_bwd_delta"}
for B, Lq, Lk in ((32, 1000, 1000), (32, 50, 1000), (32, 50, 50)):
    torch.manual_seed(0)
    q = torch.randn(B, Lq, H * dk, device=dev).to(dt).requires_grad_()
    k, v = (torch.randn(B, Lk, H * dk, device=dev).to(dt).requires_grad_() for _ in range(2))
    go = torch.randn(B, Lq, H * dk, device=dev).to(dt)
    m = F.LengthMask(torch.full((B,), Lk, dtype=torch.int64, device=dev), Lq, Lk)
    def run():
        out, _ = F.attention_core(q, k, v, m, n_head=H, dropout_p=p, seed=7)
        out.backward(go)
    for _ in range(3): run()
    torch.cuda.synchronize()
    lib.st_profile_reset(); lib.st_profile_enable(1)
    reps = 10
    for _ in range(reps): run()
    torch.cuda.synchronize(); lib.st_profile_enable(0)
    row = []
    for c, nm in names.items():
        t, w, n = C.c_double(), C.c_double(), C.c_int64()
        stb._lib.check(lib.st_profile_read(c, C.byref(t), C.byref(w), C.byref(n)))
        if n.value: row.append(f"{nm} {t.value / n.value * 1e3:7.1f} us")
    print(f"B={B} Lq={Lq} Lk={Lk} p={p} {str(dt)[6:]}: " + "  ".join(row))
