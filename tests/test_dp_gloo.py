"""World-size-2 data-parallel logic on CPU with the gloo backend (the N>1 path of bench.py minus the kernels):
flat parameter/gradient buffers, one all-reduce per step, parameter broadcast, Noam schedule.
Mirrors train_multi.py:136-139,161-163,176-177."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from speech_tranformer_pytorch_b200 import parallel as P
        torch.manual_seed(100 + rank)                      # different initial parameters on each rank
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        tr = P.DataParallelTrainer(net, d_model=512, n_warmup_steps=100, max_grad_norm=5.0)
        # parameters are views of ONE flat buffer, 16-byte aligned each
        assert tr.fp.numel == sum((p.numel() + 3) // 4 * 4 for p in net.parameters())
        for p, o in zip(tr.fp.params, tr.fp.offsets):
            assert p.data_ptr() == tr.fp.flat.data_ptr() + 4 * o and o % 4 == 0
        tr.broadcast_parameters(0)                         # train_multi.py:176
        flat0 = tr.fp.flat.clone()
        gathered = [torch.empty_like(flat0) for _ in range(world)]
        dist.all_gather(gathered, flat0)
        assert all(torch.equal(g, gathered[0]) for g in gathered), "broadcast must equalise parameters"
        # one step on a rank-dependent shard (DistributedSampler analogue: disjoint utterances per rank)
        torch.manual_seed(7)
        x_all, y_all = torch.randn(8, 6), torch.randn(8, 3)
        xs, ys = x_all[rank::world], y_all[rank::world]
        tr.zero_grad()
        loss = ((net(xs) - ys) ** 2).mean()
        loss.backward()
        for p, o in zip(tr.fp.params, tr.fp.offsets):      # autograd accumulated INTO the flat gradient buffer
            assert p.grad.data_ptr() == tr.fp.grad.data_ptr() + 4 * o
        local = tr.fp.grad.clone()
        tr.allreduce_gradients()                           # train_multi.py:161-163 — ONE collective
        summed = tr.fp.grad.clone()
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local)
        assert torch.allclose(summed, sum(parts)), "all-reduce(SUM) of the flat buffer"
        # the averaged gradient equals the gradient of the mean loss over the union of the shards
        ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
        with torch.no_grad():
            for pr, p in zip(ref.parameters(), net.parameters()):
                pr.copy_(p)
        (((ref(x_all) - y_all) ** 2).mean()).backward()
        for pr, p in zip(ref.parameters(), net.parameters()):
            assert torch.allclose(p.grad / world, pr.grad, atol=1e-6)
        q.put((rank, "ok", float(loss)))
    except Exception as e:  # surface the failure in the parent
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _worker_overlap(rank, world, port, q):
    """Bucketed, overlapped gradient exchange: buckets whose parameters were all written directly are reduced early
    (asynchronously, in backward order), the rest after backward; the result equals one all-reduce of the buffer."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from speech_tranformer_pytorch_b200 import parallel as P, functional as F
        torch.manual_seed(5)
        net = torch.nn.Sequential(*[torch.nn.Linear(16, 16) for _ in range(6)])
        tr = P.DataParallelTrainer(net, d_model=512, bucket_mb=4 * 272 / (1 << 20))      # 272 floats: one layer (weight + bias) per bucket
        assert tr.overlap and len(tr.buckets.items) >= 5
        assert tr.buckets.items[0].lo == 0 and tr.buckets.items[-1].hi == tr.fp.numel
        assert all(a.hi == b.lo for a, b in zip(tr.buckets.items, tr.buckets.items[1:])), "buckets tile the buffer"
        tr.zero_grad()
        torch.manual_seed(50 + rank)
        tr.fp.grad.copy_(torch.randn(tr.fp.numel))                     # this rank's local gradient
        local = tr.fp.grad.clone()
        # forward used every parameter once, except the first layer's weight twice (second use = autograd fallback)
        for s in tr.fp.sinks:
            s.uses = 1
        tr.fp.sinks[0].uses = 2
        # backward order = reverse registration order; each operator writes (weight, bias) then notifies
        early_before = tr.early_launches
        for i in reversed(range(0, len(tr.fp.sinks), 2)):
            F._notify(tr.fp.sinks[i:i + 2], direct=[s.view for s in tr.fp.sinks[i:i + 2]])
        started = [b.started for b in tr.buckets.items]
        assert not started[0], "a bucket holding a parameter with an outstanding use must wait for the end"
        assert all(started[1:]) and tr.early_launches - early_before == len(started) - 1
        assert tr.buckets.pending_ranges() == [(0, tr.buckets.items[0].hi)]
        tr.allreduce_gradients()
        assert all(b.work is None for b in tr.buckets.items)
        parts = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(parts, local)
        assert torch.allclose(tr.fp.grad, sum(parts), atol=1e-6), "bucketed exchange == one all-reduce(SUM)"
        # next step: bookkeeping is reset, nothing is early when no operator notifies
        tr.zero_grad()
        assert all(s.uses == 0 and s.done == 0 for s in tr.fp.sinks)
        assert tr.buckets.pending_ranges() == [(0, tr.fp.numel)]
        q.put((rank, "ok", 0.0))
    except Exception:
        import traceback
        q.put((rank, "fail", traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def _run_world2(worker):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in results:
        assert status == "ok", f"rank {rank}: {info}"


def test_overlapped_bucketed_allreduce_world2():
    _run_world2(_worker_overlap)


def test_flat_buffer_data_parallel_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    for rank, status, info in results:
        assert status == "ok", f"rank {rank}: {info}"


def test_noam_schedule_matches_reference_formula():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import numpy as np
    from speech_tranformer_pytorch_b200 import parallel as P
    for step in (1, 10, 11999, 12000, 12001, 50000):       # Optim.py:39-41
        ref = np.power(512, -0.5) * np.min([np.power(step, -0.5), np.power(12000, -1.5) * step])
        assert abs(P.noam_lr(512, 12000, step) - ref) < 1e-12


def test_grad_buckets_tile_the_buffer_for_any_bucket_size():
    """parallel.GradBuckets: contiguous, whole parameters, cover [0, numel) exactly for tiny / huge bucket sizes; the
    readiness bookkeeping only fires when every parameter of a bucket was used and written (host logic, no process group)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from speech_tranformer_pytorch_b200 import parallel as P
    net = torch.nn.Sequential(torch.nn.Linear(7, 5), torch.nn.Linear(5, 3), torch.nn.Linear(3, 2))
    tr = P.DataParallelTrainer(net, d_model=512)
    fp = tr.fp
    for floats in (1, 8, 40, 10 ** 9):
        gb = P.GradBuckets(fp, floats)
        assert gb.items[0].lo == 0 and gb.items[-1].hi == fp.numel
        assert all(a.hi == b.lo for a, b in zip(gb.items, gb.items[1:]))
        assert sum(len(b.sinks) for b in gb.items) == len(fp.sinks)
        starts = set(fp.offsets) | {fp.numel}
        assert all(b.lo in starts and b.hi in starts for b in gb.items), "buckets hold whole parameters"
        assert all(b.hi - b.lo >= floats for b in gb.items[:-1])
        assert gb.pending_ranges() == [(0, fp.numel)]
    gb = P.GradBuckets(fp, 1)                       # one parameter per bucket
    tr.zero_grad()
    s_last = fp.sinks[-1]
    assert gb.mark_done([s_last]) == []             # not used by any forward operator: never "final"
    s_last.uses, s_last.done = 2, 1
    assert gb.mark_done([s_last]) == []             # one of two uses still outstanding
    s_last.done = 2
    ready = gb.mark_done([s_last])
    assert [(b.lo, b.hi) for b in ready] == [(fp.offsets[-1], fp.numel)] and ready[0].started
    assert gb.mark_done([s_last]) == []             # a started bucket is not handed out twice
    assert gb.pending_ranges() == [(0, fp.offsets[-1])]
    gb.reset()
    assert gb.pending_ranges() == [(0, fp.numel)]


def test_eval_forward_does_not_unbalance_the_sink_counters():
    """a forward that never sees a backward (torch.no_grad(): validation, IncrementalDecoder.start) must not count
    as a use — otherwise that rank's uses/done never match and it launches different collectives than its peers."""
    import sys
    import types
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from speech_tranformer_pytorch_b200 import parallel as P, functional as F
    net = torch.nn.Linear(8, 4)
    tr = P.DataParallelTrainer(net, d_model=512)
    tr.zero_grad()
    params = (net.weight, net.bias)
    eval_ctx = types.SimpleNamespace(needs_input_grad=(False, False, False))
    assert F._sinks_of(params, eval_ctx) is None and all(s.uses == 0 for s in tr.fp.sinks)
    train_ctx = types.SimpleNamespace(needs_input_grad=(False, True, True))
    sinks = F._sinks_of(params, train_ctx)
    assert sinks is not None and all(s.uses == 1 for s in tr.fp.sinks)
    direct, zeroed = F._claim(sinks)
    assert direct is not None and zeroed == 1
    F._notify(sinks, direct)
    assert all(s.done == s.uses == 1 for s in tr.fp.sinks)


def test_overlap_refuses_gradient_accumulation():
    """with early bucket reductions a second forward before zero_grad would add local gradients on top of reduced
    sums; the sinks of an overlapping trainer refuse it (overlap=False accumulates through autograd as before)."""
    import sys
    import types
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from speech_tranformer_pytorch_b200 import parallel as P, functional as F
    net = torch.nn.Linear(8, 4)
    tr = P.DataParallelTrainer(net, d_model=512)
    for s in tr.fp.sinks:                       # what DataParallelTrainer(overlap=True) does when world > 1
        s.no_accumulate = True
    ctx = types.SimpleNamespace(needs_input_grad=(True, True, True))
    tr.zero_grad()
    sinks = F._sinks_of((net.weight, net.bias), ctx)
    F._claim(sinks)
    with pytest.raises(RuntimeError, match="accumulation"):
        F._sinks_of((net.weight, net.bias), ctx)
    tr.zero_grad()                              # a new step is fine again
    assert F._sinks_of((net.weight, net.bias), ctx) is not None


def test_trainer_state_dict_round_trip():
    """Utils.save_model / train.py:110-114 checkpoint {'model', 'optimizer'}: the optimizer half in torch.optim.Adam's
    per-parameter format, plus the global step (Noam rate, bias correction) the reference forgets."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from speech_tranformer_pytorch_b200 import parallel as P
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    tr = P.DataParallelTrainer(net, d_model=512)
    tr.exp_avg.copy_(torch.randn(tr.fp.numel))
    tr.exp_avg_sq.copy_(torch.rand(tr.fp.numel))
    tr.global_step = 1234
    sd = tr.state_dict()
    ref = torch.optim.Adam(net.parameters(), betas=(0.9, 0.98), eps=1e-9).state_dict()
    assert sd["param_groups"][0]["params"] == ref["param_groups"][0]["params"]
    assert set(sd["state"]) == set(range(4)) and all(set(v) == {"step", "exp_avg", "exp_avg_sq"} for v in sd["state"].values())
    assert sd["state"][0]["exp_avg"].shape == net[0].weight.shape
    opt = torch.optim.Adam(net.parameters(), betas=(0.9, 0.98), eps=1e-9)
    opt.load_state_dict({"state": sd["state"], "param_groups": [dict(ref["param_groups"][0])]})    # loads into stock Adam
    net2 = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Linear(5, 3))
    tr2 = P.DataParallelTrainer(net2, d_model=512)
    tr2.load_state_dict(sd)
    assert tr2.global_step == 1234
    for p_, o in zip(tr.fp.params, tr.fp.offsets):            # (the alignment padding between parameters is not state)
        n = p_.numel()
        assert torch.equal(tr2.exp_avg[o:o + n], tr.exp_avg[o:o + n]) and torch.equal(tr2.exp_avg_sq[o:o + n], tr.exp_avg_sq[o:o + n])
