"""Stress config (BASELINE.json configs[3]): ONE synthetic chromosome with N windows / ~2K+N stored entries,
row-partitioned over the ranks of a torchrun launch; full train step (both strands, 2-layer gated GCN, BCE,
backward, SGD) with NCCL all-gather of the SpMM input panels.  Rank 0 prints one JSON line.

    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_rowpartition.py [N_rows] [K_pairs]
"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from chromegcn_b200 import ops, synthetic, dist as cdist
from chromegcn_b200.chrome_models import ChromeGCN
from chromegcn_b200.graph import HiCGraph
from chromegcn_b200.optim import FlatSGD

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
k_pairs = int(float(sys.argv[2])) if len(sys.argv) > 2 else 25_000_000
steps, warmup = 5, 2
D = int(os.environ.get("CGCN_D", "128"))            # d_model: 128 (main.py:62), 256 or 512 (stress configuration)
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
a = synthetic.make_pattern_direct(n, k_pairs, seed=77)
a = (a + __import__("scipy.sparse", fromlist=["eye"]).eye(n, format="csr")).tocsr(); a.sort_indices()
parts = cdist.row_partition(n, world); b, e = parts[rank]
lp, lc = cdist.local_rows_csr(a.indptr, a.indices, b, e)
nnz_total = int(a.nnz)
g = HiCGraph.from_csr_pattern(lp, lc, dev, add_selfloops=False)
del a
torch.manual_seed(0)
m = ChromeGCN(D, D, 103, 0.2, True, 2).to(dev).train()
if world > 1:
    for p in m.parameters(): dist.broadcast(p.data, 0)
opt = FlatSGD(m, lr=0.25)
exchange = os.environ.get("CGCN_EXCHANGE", "peer")            # "peer": NVLink loads inside the SpMM ; "nccl": all-gather
step = cdist.RowPartitionedStep(m, g, parts, rank, 2, exchange=exchange)
gen = torch.Generator(device=dev).manual_seed(100 + rank)
panel = torch.randn(e - b, 2, D, device=dev, generator=gen)
tgt = (torch.rand(e - b, 103, device=dev, generator=gen) < 0.05).float()
loss = torch.zeros(1, device=dev)
def one():
    step.run(panel, tgt, loss, train=True)
    opt.step()
for _ in range(warmup): one()
if world > 1: dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps): one()
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev, dtype=torch.float64)
if world > 1: dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({"workload": "ST: one chromosome N=%d, %d stored entries, d=%d, 2 strands, row-partitioned x%d" % (n, nnz_total, D, world),
                      "n_gpus": world, "ms_per_step": float(ms.item()), "GE_per_s": nnz_total / float(ms.item()) / 1e6,
                      "exchange": exchange,
                      "allgather_bytes_per_step_per_rank": (3 * n * 2 * D * 4) if exchange == "nccl" else 0, "loss": float(loss.item() / (warmup + steps))}))
step.close()
if world > 1: dist.destroy_process_group()
