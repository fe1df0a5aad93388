"""Multi-GPU plumbing for the chromosome model: one process per GPU, `torch.distributed` (NCCL on
the GPU box, gloo in CPU tests) for the exchange steps.

The reference trains on one GPU (README.md:45) and takes one optimiser step per chromosome
(finetune.py:39-49).  Two ways the path shards:

* whole genome -- chromosomes are independent graphs that share only the 46 825 parameters, so
  they are bin-packed (LPT) onto ranks; the ranks walk their lists in lock-step "rounds", and each
  round ends with ONE all-reduce of the flat gradient buffer (~190 KB) and the same optimiser step
  on every rank.  Semantics: one step per round on the mean of that round's per-chromosome
  gradients (a grouped version of the reference's sequential steps; with world_size 1 it is
  exactly the reference's trajectory).
* one oversized graph -- contiguous row blocks per rank (`row_partition`); because the pattern is
  symmetric the same partition serves forward and backward, and the exchange step is an all-gather
  of the `[rows_local, width]` feature panel before each SpMM (`allgather_panel`).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def chromosome_cost(n: int, nnz: int, d: int = 128, strands: int = 2) -> float:
    """Bytes moved per train step, the LPT weight: 3 SpMM passes + ~40 streamed panels."""
    row = 4.0 * d * strands
    return 3.0 * nnz * (4.0 + row) + 40.0 * n * row


def lpt_shards(costs: Dict[str, float], world: int) -> List[List[str]]:
    """Longest-processing-time bin packing; deterministic (ties by name)."""
    bins: List[List[str]] = [[] for _ in range(world)]
    load = [0.0] * world
    for name in sorted(costs, key=lambda k: (-costs[k], k)):
        r = min(range(world), key=lambda i: (load[i], i))
        bins[r].append(name)
        load[r] += costs[name]
    return bins


def num_rounds(shards: Sequence[Sequence[str]]) -> int:
    return max((len(s) for s in shards), default=0)


def active_in_round(shards: Sequence[Sequence[str]], t: int) -> int:
    return sum(1 for s in shards if t < len(s))


def allreduce_gradients(flat_grad: torch.Tensor, group=None) -> None:
    """The one collective of the chromosome-sharded step: sum of the flat gradient buffer."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)


def sharded_train_epoch(engine, optimizer, shards: Sequence[Sequence[str]], rank: int, graphs, panels, targets,
                        probs: Dict[str, torch.Tensor], losses: torch.Tensor, group=None) -> None:
    """One pass over all chromosomes on `len(shards)` ranks.  `graphs/panels/targets/probs` hold this
    rank's chromosomes (device resident).  `losses[t]` receives the loss of this rank's t-th chromosome."""
    from .engine import flat_params
    mine = shards[rank]
    fp = flat_params(engine.model)
    for t in range(num_rounds(shards)):
        if t < len(mine):
            c = mine[t]
            engine.run(graphs[c], panels[c], targets[c], probs[c], losses[t: t + 1], train=True)
        else:
            fp.flat_grad.zero_()
        allreduce_gradients(fp.flat_grad, group)
        optimizer.grad_scale = 1.0 / max(active_in_round(shards, t), 1)
        optimizer.step()


def sync_batchnorm_buffers(model, group=None) -> None:
    """Average BatchNorm running statistics across ranks (each rank saw different chromosomes)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    w = dist.get_world_size(group)
    for buf in (model.batch_norm.running_mean, model.batch_norm.running_var):
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        buf.div_(w)


# ------------------------------------------------------------------ one oversized graph: row partition
def row_partition(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, nearly equal row blocks `[begin, end)` per rank."""
    base, extra = divmod(n, world)
    out, b = [], 0
    for r in range(world):
        e = b + base + (1 if r < extra else 0)
        out.append((b, e))
        b = e
    return out


def local_rows_csr(indptr: np.ndarray, indices: np.ndarray, begin: int, end: int) -> Tuple[np.ndarray, np.ndarray]:
    """CSR of rows `[begin, end)` with GLOBAL column indices (the gathered panel is global)."""
    ip = np.asarray(indptr, dtype=np.int64)
    lo, hi = ip[begin], ip[end]
    return (ip[begin: end + 1] - lo).astype(np.int32), np.asarray(indices[lo:hi], dtype=np.int32)


def allgather_panel(local: torch.Tensor, parts: Sequence[Tuple[int, int]], out: torch.Tensor = None, group=None) -> torch.Tensor:
    """Exchange step of the row-partitioned SpMM: every rank contributes its `[rows_local, width]`
    block, every rank ends with the full `[n, width]` panel.  Blocks may differ by one row."""
    world = len(parts)
    n = parts[-1][1]
    width = local.shape[1]
    if out is None:
        out = torch.empty(n, width, dtype=local.dtype, device=local.device)
    if world == 1:
        out.copy_(local)
        return out
    # blocks may differ by one row: gather fixed-size padded blocks, then drop the padding
    rows_max = max(e - b for b, e in parts)
    rank = dist.get_rank(group)
    b0, e0 = parts[rank]
    send = torch.zeros(rows_max, width, dtype=local.dtype, device=local.device)
    send[: e0 - b0] = local
    recv = torch.empty(world, rows_max, width, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv.view(world * rows_max, width), send, group=group)
    for r, (b, e) in enumerate(parts):
        out[b:e] = recv[r, : e - b]
    return out
