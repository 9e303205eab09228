"""Attention core forward + backward at the encoder shape in bf16 / fp16 inside a cudaProfiler range (ncu --profile-from-start off).
python tools/ncu_attn16.py [bf16|fp16] [dropout]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
F = stb.functional
dt = torch.float16 if (len(sys.argv) > 1 and sys.argv[1] == "fp16") else torch.bfloat16
p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
B, H, L, dk = 32, 8, 1000, 64
dev = "cuda:0"
torch.manual_seed(0)
q, k, v = (torch.randn(B, L, H * dk, device=dev).to(dt).requires_grad_() for _ in range(3))
go = torch.randn(B, L, H * dk, device=dev).to(dt)
lens = torch.full((B,), L, dtype=torch.int64, device=dev)
m = F.LengthMask(lens, L, L)
for _ in range(2):
    out, _ = F.attention_core(q, k, v, m, n_head=H, dropout_p=p, seed=7)
    out.backward(go)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out, _ = F.attention_core(q, k, v, m, n_head=H, dropout_p=p, seed=7)
out.backward(go)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
