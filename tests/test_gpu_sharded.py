"""GPU, two ranks over NCCL: the whole-genome mode of SURVEY.md 8(e).  Chromosomes are sharded over the ranks, every
lock-step round ends with ONE all-reduce of the flat gradient buffer and the same optimiser step on every rank.  The
parameters after a pass must equal the oracle's: per round, the mean of the per-chromosome gradients (finetune.py:39-48
at the round's weights, all ranks' chromosomes) applied by SGD (utils/util_methods.py:18-19); and both ranks must hold
bit-identical parameters."""
import os
import socket

import numpy as np
import pytest
import torch

from oracle import adjacency as oadj
from oracle import gcn as ogcn

pytestmark = pytest.mark.gpu
NCLASS = 7
CHROMS = ["chr18", "chr19", "chr20", "chr21", "chr22"]


def _inputs():
    from chromegcn_b200 import synthetic
    data = {}
    for i, c in enumerate(CHROMS):
        h = synthetic.make_hic(c, hic_edges=2400, n_windows=260 + 30 * i, n_bins=800)
        ip, ix = oadj.build_adjacency_numpy(h.window_starts, h.bin1, h.bin2, h.val, h.norm, 1, 2400)
        data[c] = (ip, ix, synthetic.make_features(c, ip.shape[0] - 1, 128, NCLASS))
    return data


def _initial_state():
    torch.manual_seed(5)
    return ogcn.stress_init_(ogcn.ChromeGCNOracle(128, 128, NCLASS, 0.0, True, 2)).state_dict()


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from chromegcn_b200 import dist as cdist, ops
    from chromegcn_b200.chrome_models import ChromeGCN
    from chromegcn_b200.engine import ChromosomeEngine
    from chromegcn_b200.graph import HiCGraph
    from chromegcn_b200.optim import FlatSGD
    dev = torch.device("cuda", rank)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        data = _inputs()
        costs = {c: cdist.chromosome_cost(data[c][0].shape[0] - 1, int(data[c][0][-1])) for c in CHROMS}
        schedule = cdist.balanced_schedule(costs, world)
        mine = cdist.schedule_shards(schedule, world)[rank]
        graphs = {c: HiCGraph.from_csr_pattern(data[c][0], data[c][1], dev) for c in mine}
        panels = {c: ops.interleave_strands([data[c][2]["forward"].to(dev), data[c][2]["backward"].to(dev)]) for c in mine}
        targets = {c: data[c][2]["target"].to(dev) for c in mine}
        probs = {c: torch.empty(graphs[c].n, NCLASS, device=dev) for c in mine}
        m = ChromeGCN(128, 128, NCLASS, 0.0, True, 2)
        m.load_state_dict(_initial_state())
        m = m.to(dev).train()
        m.gemm_impl = 1
        losses = torch.zeros(max(len(mine), 1), device=dev)
        cdist.sharded_train_epoch(ChromosomeEngine(m, 2), FlatSGD(m, lr=0.25), schedule, rank, graphs, panels, targets, probs, losses)
        torch.cuda.synchronize(dev)
        # numpy arrays pickle by value (tensors would travel as shared-memory handles through the manager process)
        out = {"schedule": schedule, "params": {k: v.detach().cpu().numpy().copy() for k, v in m.state_dict().items()}}
        # the library's own NCCL entry points (cgcn_comm_*, what a host without Python calls) against torch.distributed
        box = [cdist.NativeComm.unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, 0)
        comm = cdist.NativeComm(box[0], world, rank, dev)
        g = torch.Generator(device=dev).manual_seed(11 + rank)
        t = torch.randn(46825, device=dev, generator=g)
        mine_t, theirs = t.clone(), t.clone()
        comm.allreduce_sum(mine_t)
        dist.all_reduce(theirs)
        part = torch.randn(300, 256, device=dev, generator=g)
        got = torch.empty(world * 300, 256, device=dev)
        want_parts = [torch.empty_like(part) for _ in range(world)]
        comm.allgather(part, got)
        dist.all_gather(want_parts, part)
        torch.cuda.synchronize(dev)
        out["native_allreduce_equal"] = bool(torch.equal(mine_t, theirs))
        out["native_allgather_equal"] = bool(torch.equal(got, torch.cat(want_parts)))
        comm.close()
        ret[rank] = out
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_chromosome_sharded_pass_two_gpus_matches_mean_gradient_oracle():
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    ret = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    a, b = ret[0], ret[1]
    assert a["schedule"] == b["schedule"]
    for r in (a, b):
        assert r["native_allreduce_equal"] and r["native_allgather_equal"]
    for k in a["params"]:
        # replicas stay bit-identical, BatchNorm buffers included: sharded_train_epoch ends with
        # sync_batchnorm_buffers (rank average of running_mean / running_var, summed num_batches_tracked), so an eval
        # pass or a checkpoint (utils/evals.py:250-263 saves the whole state_dict) does not depend on the rank
        assert np.array_equal(a["params"][k], b["params"][k]), k
    assert int(a["params"]["batch_norm.num_batches_tracked"]) == 2 * len(CHROMS)      # one update per strand call, all ranks
    assert not np.allclose(a["params"]["batch_norm.running_mean"], 0.0)
    # oracle: per round one SGD step on the mean gradient of the round's chromosomes (both ranks')
    data = _inputs()
    om = ogcn.ChromeGCNOracle(128, 128, NCLASS, 0.0, True, 2)
    om.load_state_dict(_initial_state())
    om = om.double().train()
    oopt = ogcn.make_optimizer(om, "sgd", 0.25)
    for rnd in a["schedule"]:
        chroms = [c for cell in rnd for c in cell]
        acc = None
        for c in chroms:
            ip, ix, f = data[c]
            oopt.zero_grad()
            ogcn.chromosome_step(om, f["forward"].double(), f["backward"].double(), f["target"].double(),
                                 ogcn.coo_adjacency(ip, ix, torch.float64), None, True)
            g = [p.grad.clone() for p in om.parameters()]
            acc = g if acc is None else [x + y for x, y in zip(acc, g)]
        for p, x in zip(om.parameters(), acc):
            p.grad = x / len(chroms)
        oopt.step()
    want = om.state_dict()
    for k, v in a["params"].items():
        if "num_batches" in k or "running" in k:
            continue
        assert ogcn.max_rel(torch.from_numpy(v), want[k]) <= 2e-5, k


def test_native_comm_world_one_is_the_identity():
    """cgcn_comm_* on a single rank (the driver's one-GPU lease): NCCL loads, the communicator initialises, the
    sum over one rank and the gather of one block leave the data as it was."""
    from chromegcn_b200 import dist as cdist
    dev = torch.device("cuda", 0)
    comm = cdist.NativeComm(cdist.NativeComm.unique_id(), 1, 0, dev)
    t = torch.randn(1000, device=dev)
    u = t.clone()
    comm.allreduce_sum(u)
    got = torch.empty_like(t)
    comm.allgather(t, got)
    torch.cuda.synchronize(dev)
    assert torch.equal(u, t) and torch.equal(got, t)
    comm.close()
