"""Cycles per tcgen05.mma.kind::tf32 (M=128, K=8) by issue style, operand form, N and number of accumulators."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import speech_tranformer_pytorch_b200 as stb
lib = stb._lib.load(); torch.zeros(1, device="cuda")
names = {0: "A smem, B K-major ", 1: "A TMEM, B K-major ", 2: "A smem, B MN-major", 3: "A TMEM, B MN-major"}
for style in (0, 16):
    for nacc in (1, 2):
        for v in range(4):
            row = []
            for n in (64, 128, 256):
                if nacc * n > 256: continue
                out = C.c_double()
                stb._lib.check(lib.st_debug_mma_bench(v | ((nacc - 1) << 2) | style, n, 200, C.byref(out)))
                row.append(f"N={n}: {out.value:6.1f} (floor {n/2:.0f})")
            print("elect.sync warp" if style else "tid==0 thread  ", f"acc={nacc}", names[v], " | ".join(row))
