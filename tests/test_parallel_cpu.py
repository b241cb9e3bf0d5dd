"""CPU, world_size 2, gloo: the data-parallel host logic (sharding + flat-bucket gradient averaging)
reproduces the single-process full-batch gradients of the oracle — the only multi-GPU behaviour the
reference defines.  The CUDA kernels are not involved (no GPU here); the same
``vit2spn.parallel`` functions are what bench.py uses under NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import vit2spn
        from vit2spn import parallel
        from oracle import vit2spn_oracle as orc
        torch.set_num_threads(2)
        state = orc.init_state(11, 0.01)
        x1, x2 = orc.synthetic_views(4, seed=5)
        a, b = parallel.shard_batch(x1, rank, world), parallel.shard_batch(x2, rank, world)
        loss, _, _, grads = orc.loss_and_grads(dict(state), a, b, 1)
        # lay the rank-local gradients out in flat buckets exactly as the model does
        names = orc.trainable_names()
        heads = [n for n in names if "head" in n]
        o2 = [n for n in names if n.startswith("online_network_2")]
        o1 = [n for n in names if n.startswith("online_network_1")]
        buckets = [torch.cat([grads[n].flatten() for n in grp]) for grp in (heads, o2, o1)]
        w = parallel.allreduce_buckets(buckets, average=True)
        assert w == world
        lsum = loss.clone()
        dist.all_reduce(lsum)
        if rank == 0:
            q.put((lsum.item() / world, [bk.clone() for bk in buckets], (heads, o2, o1)))
        # parameter broadcast keeps replicas identical
        model = vit2spn.DualStreamNetwork()
        parallel.broadcast_parameters(model)
        chk = model._head_store.flat.double().sum()
        lst = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(lst, chk)
        assert all(torch.equal(lst[0], t) for t in lst)
        with pytest.raises(ValueError):
            parallel.shard_batch(torch.zeros(5, 1), rank, world)
        # the overlapped, range-wise all-reduce (what bench.py uses at N > 1): put the rank-local oracle gradients into
        # the model's flat gradient buffers, drive the hooks in the order ssp_step does, and compare with the average
        gmap = dict(model.named_parameters())
        for s_ in model._stores()[:2] + [model._head_store]:
            s_.grads()
        with torch.no_grad():
            for n in names:
                gmap[n].grad.copy_(grads[n])
        sync = parallel.OverlappedGradSync(model, splits=(9, 5, 2), comm_sms=0)
        assert sync.ranges == [(12, 9), (9, 5), (5, 2), (2, 0)]
        sync.begin()
        sync.heads_ready()
        for hi, lo in sync.ranges:
            sync.range_ready(hi, lo)
        assert sync.finish(optimizer=None) == world              # no optimizer given: the mean is applied in place
        if rank == 0:
            q.put({n: gmap[n].grad.numpy().copy() for n in names})     # by value (numpy), not via shared memory
        covered = torch.zeros(model._stores()[0].active_numel, dtype=torch.int32)
        for hi, lo in sync.ranges:
            sl = sync._slice(model._stores()[0], hi, lo)
            off = (sl.data_ptr() - model._stores()[0].flat_grad.data_ptr()) // 4
            covered[off:off + sl.numel()] += 1
        assert bool((covered == 1).all()), "every active gradient element must be reduced exactly once"
    finally:
        dist.destroy_process_group()


def test_two_rank_gradient_average_equals_full_batch():
    from oracle import vit2spn_oracle as orc
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    mean_loss, buckets, groups = q.get()
    overlapped = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    state = orc.init_state(11, 0.01)
    x1, x2 = orc.synthetic_views(4, seed=5)
    loss, _, _, grads = orc.loss_and_grads(dict(state), x1, x2, 1)
    assert abs(loss.item() - mean_loss) < 1e-7                      # sharded mean == global mean (D3)
    for bk, grp in zip(buckets, groups):
        ref = torch.cat([grads[n].flatten() for n in grp])
        rel = float((bk - ref).norm() / ref.norm())
        assert rel < 1e-5, rel
    num = sum(float(((torch.from_numpy(overlapped[n]) - grads[n]) ** 2).sum()) for n in grads)
    den = sum(float((grads[n] ** 2).sum()) for n in grads)
    assert (num / den) ** 0.5 < 1e-5
