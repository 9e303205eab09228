"""GPU parity of incremental decoding with K/V reuse and on-device beam search against the
full-prefix decoder and oracle/decode_port.py (the reference's Decode.py / Beam.py algorithm on the CPU oracle)."""
import pytest
import torch

from helpers import golden, relerr, t
from oracle import decode_port, model_port

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
TOL_MODEL = 3e-3          # composed-model tolerance, see tests/test_gpu_model.py
CFG = dict(d_model=64, n_heads=2, num_enc_layer=2, num_dec_layer=2, vocab_size=31)


@pytest.fixture(scope="module")
def stb():
    import speech_tranformer_pytorch_b200 as m
    m.build()
    m._lib.check(m._lib.load().st_device_check(0))
    return m


def _model(stb, g, sharpen=1.0):
    from test_gpu_model import _small_model
    net = _small_model(stb, g)
    P = {k[2:]: t(v).double() for k, v in g.items() if k.startswith("p.") and not k.endswith(".pe")}
    if sharpen != 1.0:   # peaky output distributions: top-k decisions far from ties, robust to TF32 vs fp64
        with torch.no_grad():
            net.tgt_word_proj.weight.mul_(sharpen)
        P["tgt_word_proj.weight"] = P["tgt_word_proj.weight"] * sharpen
    return net, P


def test_incremental_steps_match_full_prefix_decoder(stb):
    """Teacher-forced: the logits of every incremental step (cached K/V, one new position) equal the last position of
    the full-prefix decoder — this library's own full forward and the CPU oracle."""
    from speech_tranformer_pytorch_b200.decode import IncrementalDecoder
    g = golden("transformer_small")
    net, P = _model(stb, g)
    inputs, in_len, targets = t(g["inputs"]), t(g["in_len"]), t(g["targets"])
    B, L = targets.shape
    beam = 2                                                     # two identical beams per utterance: exercises the beam-as-query-rows layout
    dec = IncrementalDecoder(net, max_len=16)
    dec.start(inputs.to(DEV), in_len.to(DEV), beam=beam)
    rep = targets.repeat_interleave(beam, 0)
    full_len = torch.full((B,), L, dtype=torch.int64)
    with torch.no_grad():
        full, _ = net(inputs.to(DEV), in_len.to(DEV), targets.to(DEV), full_len.to(DEV))      # causal: position t sees 0..t
    ref = model_port.forward(P, CFG, inputs, in_len, targets, full_len)
    for step in range(L):
        logits = dec.step(rep[:, step].to(DEV))
        assert logits.shape == (B * beam, 31)
        assert torch.equal(logits[0::beam], logits[1::beam])     # identical beams -> identical rows, bit for bit
        assert relerr(logits[0::beam], ref[:, step]) < TOL_MODEL, step
        assert relerr(logits[0::beam], full[:, step]) < TOL_MODEL, step


@pytest.mark.parametrize("beam,eos", [(1, 2), (4, 2), (4, 3)])   # eos = 3: some utterances finish early (frozen beams)
def test_beam_search_matches_oracle(stb, beam, eos):
    from speech_tranformer_pytorch_b200.decode import beam_search
    g = golden("transformer_small")
    net, P = _model(stb, g, sharpen=12.0)
    inputs, in_len = t(g["inputs"]), t(g["in_len"])
    hyps, scores = beam_search(net, inputs.to(DEV), in_len.to(DEV), beam=beam, max_len=10, n_best=min(beam, 2), eos=eos)
    rhyps, rscores = decode_port.beam_search(P, CFG, inputs.double(), in_len, beam=beam, max_len=10, n_best=min(beam, 2), eos=eos)
    assert relerr(scores, rscores) < 2e-2                         # sums of up to 10 log-probabilities of sharpened logits
    for b in range(inputs.size(0)):
        for k in range(len(rhyps[b])):
            if hyps[b][k] != rhyps[b][k]:                        # only acceptable at a near-tie between hypotheses
                assert abs(float(scores[b, k]) - float(rscores[b, k])) < 1e-2 * max(1.0, abs(float(rscores[b, k]))), (b, k, hyps[b][k], rhyps[b][k])
    assert sum(hyps[b][0] == rhyps[b][0] for b in range(inputs.size(0))) >= inputs.size(0) - 1


def test_decode_requires_eval_mode(stb):
    from speech_tranformer_pytorch_b200.decode import IncrementalDecoder
    g = golden("transformer_small")
    net, _ = _model(stb, g)
    net.train()
    with pytest.raises(RuntimeError):
        IncrementalDecoder(net).start(t(g["inputs"]).to(DEV), t(g["in_len"]).to(DEV))


def test_cuda_graph_replay_is_bit_identical(stb):
    """A persistent decoder captures each position into a CUDA graph the second time it sees a shape; results of the
    eager pass, the capturing pass and the replaying pass must be identical, including after new inputs."""
    from speech_tranformer_pytorch_b200.decode import IncrementalDecoder, beam_search
    g = golden("transformer_small")
    net, _ = _model(stb, g, sharpen=12.0)
    inputs, in_len = t(g["inputs"]).to(DEV), t(g["in_len"]).to(DEV)
    dec = IncrementalDecoder(net, max_len=10, use_graphs=True)
    runs = [beam_search(net, inputs, in_len, beam=4, max_len=10, n_best=2, eos=3, decoder=dec) for _ in range(3)]
    assert len(dec._graphs) > 0
    for hyps, scores in runs[1:]:
        assert hyps == runs[0][0] and torch.equal(scores, runs[0][1])
    other = inputs.flip(0).contiguous()                                   # same shape, different utterances: graphs are reused
    want = beam_search(net, other, in_len.flip(0).contiguous(), beam=4, max_len=10, n_best=2, eos=3)
    got = beam_search(net, other, in_len.flip(0).contiguous(), beam=4, max_len=10, n_best=2, eos=3, decoder=dec)
    assert got[0] == want[0] and torch.equal(got[1], want[1])


@pytest.mark.parametrize("n,H,dk,t", [(6, 2, 32, 0), (20, 8, 64, 17), (3, 4, 128, 49)])
def test_decode_self_attention_kernel(stb, n, H, dk, t):
    """st_decode_self_attn (one new query per hypothesis over the time-major K/V cache) against softmax(q K^T / sqrt(dk)) V
    in float64, for every supported head width; also checks that the new K/V rows were appended at position t."""
    import ctypes as C
    lib = stb._lib.load()
    d = H * dk
    g = torch.Generator().manual_seed(n + t)
    L_max = 50
    kc = torch.randn(L_max, n, d, generator=g)
    vc = torch.randn(L_max, n, d, generator=g)
    qkv = torch.randn(n, 3 * d, generator=g)
    kcd, vcd, qd = kc.to(DEV), vc.to(DEV), qkv.to(DEV)
    ctx = torch.empty(n, d, device=DEV)
    stb._lib.check(lib.st_decode_self_attn(qd.data_ptr(), kcd.data_ptr(), vcd.data_ptr(), t, n, H, dk, ctx.data_ptr(), 0, None,
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert torch.equal(kcd[t].cpu(), qkv[:, d:2 * d]) and torch.equal(vcd[t].cpu(), qkv[:, 2 * d:])
    assert torch.equal(kcd[:t].cpu(), kc[:t]) and torch.equal(vcd[t + 1:].cpu(), vc[t + 1:])      # nothing else touched
    K = torch.cat([kc[:t], qkv[None, :, d:2 * d]], 0).double().view(t + 1, n, H, dk)
    V = torch.cat([vc[:t], qkv[None, :, 2 * d:]], 0).double().view(t + 1, n, H, dk)
    q = qkv[:, :d].double().view(n, H, dk)
    s = torch.einsum("nhd,tnhd->nht", q, K) / dk ** 0.5
    ref = torch.einsum("nht,tnhd->nhd", torch.softmax(s, -1), V).reshape(n, d)
    assert relerr(ctx, ref) < 1e-5
    # with a slot table: position j of hypothesis i's history is read from slot perm[j][i] (beam-search re-parenting
    # without moving the caches); the appended row goes to the hypothesis' own slot and is recorded in the table
    perm = torch.stack([torch.randperm(n, generator=g) for _ in range(L_max)]).to(torch.int32)
    slot = perm.to(DEV)
    kcd, vcd = kc.to(DEV), vc.to(DEV)
    stb._lib.check(lib.st_decode_self_attn(qd.data_ptr(), kcd.data_ptr(), vcd.data_ptr(), t, n, H, dk, ctx.data_ptr(), 0,
                                           slot.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert torch.equal(slot[t].cpu(), torch.arange(n, dtype=torch.int32)) and torch.equal(slot[:t].cpu(), perm[:t])
    idx = perm[:t].long()
    Kp = torch.cat([torch.gather(kc[:t], 1, idx[:, :, None].expand(-1, -1, d)), qkv[None, :, d:2 * d]], 0).double().view(t + 1, n, H, dk)
    Vp = torch.cat([torch.gather(vc[:t], 1, idx[:, :, None].expand(-1, -1, d)), qkv[None, :, 2 * d:]], 0).double().view(t + 1, n, H, dk)
    s = torch.einsum("nhd,tnhd->nht", q, Kp) / dk ** 0.5
    ref = torch.einsum("nht,tnhd->nhd", torch.softmax(s, -1), Vp).reshape(n, d)
    assert relerr(ctx, ref) < 1e-5


@pytest.mark.parametrize("B,beam,V,first", [(5, 10, 4337, 0), (3, 4, 31, 0), (4, 10, 4337, 1), (2, 1, 57, 0), (2, 32, 101, 0)])
def test_beam_step_kernel_matches_torch_bookkeeping(stb, B, beam, V, first):
    """st_beam_step against the tensor formulation of Beam.advance (Beam.py:43-74): scores within 1e-5, back-pointers /
    symbols / done flags / re-parenting vectors exactly, including frozen (finished) utterances and the first position."""
    import ctypes as C
    lib = stb._lib.load()
    g = torch.Generator().manual_seed(B * 100 + beam)
    ld = (V + 3) // 4 * 4
    store = torch.randn(B * beam, ld, generator=g).mul_(3).to(DEV)
    logits = store[:, :V]
    scores0 = (torch.randn(B, beam, generator=g).abs().neg() if not first else torch.zeros(B, beam)).to(DEV)
    done0 = torch.zeros(B, dtype=torch.bool)
    done0[B // 2] = True
    done0 = done0.to(DEV)
    EOS_, PAD_ = 3, 0
    if not first:   # make utterance 0's best continuation EOS
        logits[0 * beam + int(scores0[0].argmax()), EOS_] = 50.0
    # reference (decode.beam_search before the kernel existed)
    logp = torch.log_softmax(logits.double(), dim=-1).view(B, beam, V)
    cand = logp + scores0.double().unsqueeze(2) if not first else logp[:, :1]
    best, idx = cand.reshape(B, -1).topk(beam, dim=1)
    pk = torch.div(idx, V, rounding_mode="floor")
    y = idx - pk * V
    keep = done0.unsqueeze(1)
    pk = torch.where(keep, torch.arange(beam, device=DEV).expand(B, -1), pk)
    y = torch.where(keep, torch.full_like(y, PAD_), y)
    want_scores = torch.where(keep, scores0.double(), best)
    want_done = done0 | (y[:, 0] == EOS_)
    # kernel
    scores, done = scores0.clone(), done0.clone()
    prev_k = torch.empty(B, beam, dtype=torch.int64, device=DEV)
    next_y = torch.empty_like(prev_k)
    parent = torch.empty(B * beam, dtype=torch.int64, device=DEV)
    tokens = torch.empty_like(parent)
    p = lambda t: C.c_void_p(t.data_ptr())
    stb._lib.check(lib.st_beam_step(p(logits), ld, B, beam, V, first, EOS_, PAD_, p(scores), p(done), p(prev_k), p(next_y),
                                    p(parent), p(tokens), None))
    torch.cuda.synchronize()
    assert torch.equal(prev_k, pk) and torch.equal(next_y, y)
    assert torch.equal(done, want_done) and (first or bool(done[0]))
    assert (scores.double() - want_scores).abs().max() < 1e-5
    base = (torch.arange(B, device=DEV) * beam).unsqueeze(1)
    assert torch.equal(parent, (base + pk).reshape(-1)) and torch.equal(tokens, y.reshape(-1))
