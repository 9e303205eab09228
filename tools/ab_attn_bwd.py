"""A/B the attention-backward kernel variants (st_set_option) on the encoder self-attention shape and the decoder
cross-attention shape: per-kernel-class time from the library's event profiler + max difference of the gradients.
python tools/ab_attn_bwd.py"""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import speech_tranformer_pytorch_b200 as stb
F = stb.functional
lib = stb._lib.load()
dev = "cuda:0"
torch.manual_seed(0)


def run(B, H, Lq, Lk, dk, opts, p=0.1):
    for k, v in opts.items():
        stb._lib.check(lib.st_set_option(k.encode(), v))
    d = H * dk
    g = torch.Generator(device=dev).manual_seed(1)
    q = torch.randn(B, Lq, d, device=dev, generator=g).requires_grad_()
    k = torch.randn(B, Lk, d, device=dev, generator=g).requires_grad_()
    v = torch.randn(B, Lk, d, device=dev, generator=g).requires_grad_()
    lens = torch.full((B,), Lk, device=dev); lens[1::2] = Lk - 37
    mask = (torch.arange(Lk, device=dev)[None, :] >= lens[:, None])[:, None, :].expand(-1, Lq, -1)
    go = torch.randn(B, Lq, d, device=dev, generator=g)
    out, _ = F.attention_core(q, k, v, mask, n_head=H, dropout_p=p, seed=1234)
    for _ in range(2):
        out.backward(go, retain_graph=True)
    q.grad = k.grad = v.grad = None
    lib.st_profile_reset(); lib.st_profile_enable(1)
    reps = 5
    for _ in range(reps):
        q.grad = k.grad = v.grad = None
        out.backward(go, retain_graph=True)
    torch.cuda.synchronize(); lib.st_profile_enable(0)
    res = {}
    for c in range(lib.st_profile_classes()):
        t, w, n = C.c_double(), C.c_double(), C.c_int64()
        lib.st_profile_read(c, C.byref(t), C.byref(w), C.byref(n))
        if n.value:
            res[lib.st_profile_class_name(c).decode()] = t.value / n.value * 1e3
    for k_ in opts:
        lib.st_set_option(k_.encode(), 0)
    return res, (q.grad.clone(), k.grad.clone(), v.grad.clone())


for name, shape in (("enc self 32x8x1000x1000", (32, 8, 1000, 1000, 64)), ("dec cross 32x8x50x1000", (32, 8, 50, 1000, 64)),
                    ("dec self 32x8x50x50", (32, 8, 50, 50, 64)), ("dk32 8x2x300x300", (8, 2, 300, 300, 32))):
    base, gb = run(*shape, {})
    print(f"{name}: base {({k: round(v, 1) for k, v in base.items() if k.startswith('attn')})} us")
    for opts in ({"attn_dq_res_smem": 1}, {"attn_dkv_res_smem": 1}, {"attn_dkv_res_smem": 2}, {"attn_dkv_no_small": 1},
                 {"attn_dkv_small_split": 1}, {"attn_dkv_small_split": 2}, {"attn_dkv_small_split": 8}):
        r, g = run(*shape, opts)
        diff = max(float((a - b).abs().max() / b.abs().max()) for a, b in zip(g, gb))
        print(f"   {opts}: {({k: round(v, 1) for k, v in r.items() if k.startswith('attn')})} us  max rel diff vs base {diff:.2e}")
