"""CPU (gloo, world_size 2): the host-side logic of the two sharding modes.
The kernels themselves need a GPU; here the exchange steps run on stand-in tensors."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from chromegcn_b200 import dist as cdist
from chromegcn_b200 import synthetic


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_lpt_shards_cover_and_balance():
    chroms = synthetic.WHOLE_GENOME
    costs = {c: cdist.chromosome_cost(synthetic.num_windows(c), 500000 + synthetic.num_windows(c)) for c in chroms}
    for world in (1, 2, 4, 8):
        shards = cdist.lpt_shards(costs, world)
        assert sorted(sum(shards, [])) == sorted(chroms)
        loads = [sum(costs[c] for c in s) for s in shards]
        assert max(loads) <= 1.25 * (sum(loads) / world) or world == 8     # chr1 alone bounds the 8-way balance
        assert cdist.lpt_shards(costs, world) == shards                    # deterministic
    s8 = cdist.lpt_shards(costs, 8)
    assert cdist.num_rounds(s8) == max(len(s) for s in s8)
    assert sum(cdist.active_in_round(s8, t) for t in range(cdist.num_rounds(s8))) == len(chroms)


def test_balanced_schedule_properties():
    chroms = synthetic.WHOLE_GENOME
    costs = {c: cdist.chromosome_cost(synthetic.num_windows(c), 500000 + synthetic.num_windows(c)) for c in chroms}
    total = sum(costs.values())
    for world in (1, 2, 4, 8):
        sch = cdist.balanced_schedule(costs, world)
        assert len(sch) == cdist.default_rounds(len(chroms), world)
        assert all(len(r) == world for r in sch)
        assert all(any(len(cell) for cell in r) for r in sch)                         # no empty round
        assert sorted(c for r in sch for cell in r for c in cell) == sorted(chroms)   # every chromosome exactly once
        eff = total / world / cdist.schedule_cost(sch, costs)
        assert eff >= {1: 0.999, 2: 0.95, 4: 0.90, 8: 0.87}[world], (world, eff)
        assert cdist.balanced_schedule(costs, world) == sch                           # deterministic
        shards = cdist.schedule_shards(sch, world)
        assert sorted(sum(shards, [])) == sorted(chroms)
    one = cdist.balanced_schedule(costs, 1)
    assert [len(r[0]) for r in one] == [1] * 23                                        # one step per chromosome on one rank


def test_row_partition_and_local_csr():
    n = 1003
    parts = cdist.row_partition(n, 8)
    assert parts[0][0] == 0 and parts[-1][1] == n and all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
    assert max(e - b for b, e in parts) - min(e - b for b, e in parts) <= 1
    a = synthetic.make_pattern_direct(n, 4000, seed=3, max_dist=50)
    for b, e in parts:
        ip, ix = cdist.local_rows_csr(a.indptr, a.indices, b, e)
        assert ip[0] == 0 and ip[-1] == ix.shape[0] == a.indptr[e] - a.indptr[b]
        assert np.array_equal(ix, a.indices[a.indptr[b]: a.indptr[e]])


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) chromosome sharding: the all-reduced flat gradient equals the sum of per-rank gradients
        g = torch.full((1000,), float(rank + 1))
        cdist.allreduce_gradients(g)
        assert torch.equal(g, torch.full((1000,), 3.0))
        # (2) row partition: all-gather of unequal row blocks rebuilds the global panel, and a mean
        #     aggregation over local rows with global columns equals the single-process result
        n, width = 101, 8
        parts = cdist.row_partition(n, world)
        gen = torch.Generator().manual_seed(0)
        full = torch.randn(n, width, generator=gen)
        b, e = parts[rank]
        got = cdist.allgather_panel(full[b:e].clone(), parts)
        assert torch.equal(got, full)
        a = synthetic.make_pattern_direct(n, 300, seed=1, max_dist=20)
        from oracle import adjacency as oadj
        rp, ci = oadj.pattern_with_selfloops(a.indptr, a.indices)
        lp, lc = cdist.local_rows_csr(rp, ci, b, e)
        local = torch.stack([got[lc[lp[i]: lp[i + 1]]].mean(0) for i in range(e - b)])
        want = torch.stack([full[ci[rp[i]: rp[i + 1]]].mean(0) for i in range(b, e)])
        assert torch.allclose(local, want)
        # (3) BatchNorm buffers after a sharded pass: rank average of the running statistics, num_batches_tracked =
        #     start + the calls of ALL ranks (what one model walking every chromosome would count)
        holder = type("M", (), {})()
        holder.batch_norm = torch.nn.BatchNorm1d(4)
        bn = holder.batch_norm
        before = bn.num_batches_tracked.clone()
        bn.running_mean.fill_(float(rank + 1))
        bn.running_var.fill_(float(10 * (rank + 1)))
        bn.num_batches_tracked.add_(3 + rank)                     # rank 0 saw 3 strand calls, rank 1 saw 4
        cdist.sync_batchnorm_buffers(holder, None, before)
        assert torch.equal(bn.running_mean, torch.full((4,), 1.5)) and torch.equal(bn.running_var, torch.full((4,), 15.0))
        assert int(bn.num_batches_tracked) == 7
        # (4) one optimiser step on the MEAN of a round's summed gradients with a stock torch optimiser (no grad_scale
        #     attribute): the buffer is scaled in place, nothing persists on the optimiser
        p = torch.nn.Parameter(torch.ones(5))
        p.grad = torch.full((5,), 4.0)
        opt = torch.optim.SGD([p], lr=0.5)
        cdist.step_on_mean(opt, p.grad, 0.25)
        assert torch.allclose(p.detach(), torch.full((5,), 0.5)) and not hasattr(opt, "grad_scale")
        ret[rank] = 1
    finally:
        dist.destroy_process_group()


def test_gloo_world2_exchange_steps():
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, ret)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(ret.keys()) == [0, 1]
