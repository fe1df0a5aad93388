#!/usr/bin/env python
"""bench.py -- ChromeGCN chromosome-model hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload wg|c1]

Metric (BASELINE.json): "GCN train step edges/sec (GE/s)".  A step is one pass of the hot path over the
whole-genome synthetic workload: for each of the 23 GM12878-shaped chromosome graphs, both strands through
the 2-layer gated GCN, BCE loss, backward, optimiser step (finetune.py:29-49).  Unit of work: one directed
stored entry of A_hat = D^-1 bin(A+I), so  value = sum_c nnz(A_hat_c) / t_step / 1e9.

  value      inputs (graphs, feature panels, targets) resident in HBM when the timed region starts;
  e2e        the same metric through the drop-in `finetune()` with pinned HOST features: H2D of every
             chromosome's features/targets and D2H of the predictions inside the timed region;
  roofline   the SpMM kernel timed alone with CUDA events (one launch per chromosome, strand-batched
             width 256) against MEASURED_PEAKS.json's copy bandwidth, algorithmic bytes
             nnz*(4+4W) + 4(N+1) + 4WN per launch;
  cpu_baseline / --impl reference   the oracle port of the reference's PyTorch CPU path (oracle/gcn.py:
             same torch.mm / torch.spmm / scipy process_graph per chromosome per epoch) on the host cores,
             on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "GCN train step edges/sec (GE/s)"
UNIT = "GE/s"
D, NCLASS, LAYERS, HIC_EDGES = 128, 103, 2, 500000
DROPOUT, LR = 0.2, 0.25      # README.md:45 recipe: SGD lr 0.25, gcn_dropout 0.2
CPU_SAMPLE = ["chr19", "chr20", "chr21", "chr22"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="wg", choices=["wg", "c1"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--gemm-impl", type=int, default=0)
    return ap.parse_args()


def workload_chroms(name):
    from chromegcn_b200 import synthetic
    return ["chr22"] if name == "c1" else list(synthetic.WHOLE_GENOME)


def workload_name(name):
    return ("C1: one chr22-sized graph (N=20000, hicsize 500000)" if name == "c1" else
            "WG: whole-genome synthetic GM12878-shaped, 23 chromosome graphs (sum N ~1.18M, hicsize 500000 each)")


# ---------------------------------------------------------------------------------------- CPU (oracle port)
def cpu_reference_run(chroms, steps, warmup, threads=None):
    """The reference's CPU path restated (oracle/gcn.py), timed on the host: per step one pass over `chroms`
    exactly as finetune.py does it (process_graph on the host for every chromosome every pass, two
    forward calls, BCE, backward, SGD)."""
    import torch
    from chromegcn_b200 import synthetic
    from oracle import adjacency as oadj
    from oracle import gcn as ogcn
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    graphs, feats, edges = {}, {}, 0
    for c in chroms:
        h = synthetic.make_hic(c, hic_edges=HIC_EDGES)
        ip, ix = oadj.build_adjacency_numpy(h.window_starts, h.bin1, h.bin2, h.val, h.norm, 1, HIC_EDGES)
        graphs[c] = (ip, ix)
        feats[c] = synthetic.make_features(c, ip.shape[0] - 1, D, NCLASS)
        edges += int(ip[-1]) + ip.shape[0] - 1
    torch.manual_seed(0)
    model = ogcn.ChromeGCNOracle(D, D, NCLASS, DROPOUT, True, LAYERS)
    opt = ogcn.make_optimizer(model, "sgd", LR)
    for _ in range(warmup):
        ogcn.finetune_epoch(model, feats, graphs, opt, "train")
    t0 = time.perf_counter()
    for _ in range(steps):
        ogcn.finetune_epoch(model, feats, graphs, opt, "train")
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return edges / dt / 1e9, dt, edges, threads


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(args.steps, 1), max(args.warmup, 0)
    chroms = ["chr22"] if args.workload == "c1" else CPU_SAMPLE
    value, dt, edges, threads = cpu_reference_run(chroms, steps, warmup)
    sample = "one pass over %s (%d stored entries) per step, oracle port of the reference PyTorch CPU path" % (
        "+".join(chroms), edges)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(args.workload), "sample": sample, "d_model": D, "gcn_layers": LAYERS,
                       "nclass": NCLASS, "optim": "sgd"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.t0 = self.t1 = None

    def mark_begin(self):
        self.t0 = time.monotonic()

    def mark_end(self):
        self.t1 = time.monotonic()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.monotonic(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        # nvidia-smi needs ~1 s to come up, so it is started before the data is built and its rows are
        # time-stamped: the reported samples are the ones that fall inside the timed region (plus one
        # sampling period of slack either side, the loop period being 20 ms).
        rows = self.rows
        window = "timed region"
        if self.t0 is not None and self.t1 is not None:
            rows = [(t, r) for t, r in self.rows if self.t0 - 0.02 <= t <= self.t1 + 0.02]
            if not rows:      # region shorter than one sampling period: fall back to the warm-up + timed load window
                rows = [(t, r) for t, r in self.rows if self.load0 - 0.02 <= t <= self.t1 + 0.02]
                window = "warm-up + timed region (timed region shorter than one 20 ms sample)"
        sm, mx, reasons = [], [], set()
        for _, r in rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ---------------------------------------------------------------------------------------- GPU arm
def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return
    import numpy as np
    import torch
    import torch.distributed as dist
    from chromegcn_b200 import _lib, ops, synthetic
    from chromegcn_b200 import dist as cdist
    from chromegcn_b200 import finetune as ft
    from chromegcn_b200.chrome_models import ChromeGCN
    from chromegcn_b200.engine import ChromosomeEngine, flat_params
    from chromegcn_b200.graph import HiCGraph
    from chromegcn_b200.optim import FlatSGD

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    steps, warmup = max(args.steps, 1), max(args.warmup, 3)
    _lib.require_cuda()
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    from chromegcn_b200 import hostbind
    binding = hostbind.bind_to_gpu(local_rank)      # before any pinned allocation (first touch decides the NUMA node)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()            # started here so that nvidia-smi is already looping when the timed region begins
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    chroms = workload_chroms(args.workload)
    # ---- synthetic inputs; adjacency built by the product path (cgcn_adj_build) for this rank's chromosomes
    sizes = {c: synthetic.num_windows(c) for c in chroms}
    costs = {c: cdist.chromosome_cost(sizes[c], HIC_EDGES + sizes[c]) for c in chroms}
    schedule = cdist.balanced_schedule(costs, world)
    shards = cdist.schedule_shards(schedule, world)
    mine = shards[rank]
    graphs, feats_host, panels, targets, probs = {}, {}, {}, {}, {}
    local_edges = 0
    for c in mine:
        h = synthetic.make_hic(c, hic_edges=HIC_EDGES)
        ip, ix = ops.adjacency_build(h.window_starts, h.bin1, h.bin2, h.val, h.norm, 1, HIC_EDGES, dev)
        graphs[c] = HiCGraph.from_csr_pattern(ip, ix, dev, name=c)
        n = graphs[c].n
        f = synthetic.make_features(c, n, D, NCLASS)
        feats_host[c] = {k: v.pin_memory() for k, v in f.items()}
        panels[c] = ops.interleave_strands([f["forward"].to(dev), f["backward"].to(dev)])
        targets[c] = ops.pack_targets(f["target"]).to(dev)      # label bit rows, the form finetune() feeds the loss kernel
        probs[c] = torch.empty(n, NCLASS, device=dev)
        local_edges += graphs[c].nnz
    edges_t = torch.tensor([local_edges], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(edges_t)
    total_edges = float(edges_t.item())

    torch.manual_seed(0)
    model = ChromeGCN(D, D, NCLASS, DROPOUT, True, LAYERS).to(dev)
    model.gemm_impl = args.gemm_impl
    model.train()
    if world > 1:
        for p in model.parameters():
            dist.broadcast(p.data, 0)
    engine = ChromosomeEngine(model, 2)
    optimizer = FlatSGD(model, lr=LR)
    losses = torch.zeros(max(len(mine), 1), device=dev)

    def one_step():
        cdist.sharded_train_epoch(engine, optimizer, schedule, rank, graphs, panels, targets, probs, losses)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler.load0 = time.monotonic()
    for _ in range(warmup):
        one_step()
    barrier()
    launches0 = _lib.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.mark_begin()
    ev0.record()
    torch.cuda.nvtx.range_push("timed")
    for _ in range(steps):
        one_step()
    torch.cuda.nvtx.range_pop()
    ev1.record()
    barrier()
    sampler.mark_end()
    ms = ev0.elapsed_time(ev1)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    lt = torch.tensor([launches], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
    ms_per_step = float(t.item()) / steps
    value = total_edges / (ms_per_step * 1e-3) / 1e9
    final_loss = float(losses.sum().item())

    # ---- e2e: drop-in finetune() with pinned host features (rank-local chromosomes), H2D/D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        import pickle
        import tempfile
        from scipy import sparse
        tmp = tempfile.mkdtemp(prefix="cgcn_bench_")
        gdict = {}
        for c in mine:
            rp, ci = graphs[c].csr_numpy()
            n = graphs[c].n
            rows = np.repeat(np.arange(n), np.diff(rp))
            keep = rows != ci                                  # the pickles hold A without self loops
            gdict[c] = sparse.csr_matrix((np.ones(int(keep.sum())), (rows[keep], ci[keep])), shape=(n, n))
        with open(os.path.join(tmp, "train_graphs_%d_SQRTVCnorm.pkl" % HIC_EDGES), "wb") as fp:
            pickle.dump(gdict, fp)
        opt_ns = argparse.Namespace(adj_type="hic", graph_root=tmp, hicsize=str(HIC_EDGES), hicnorm="SQRTVC")
        # per pass: both strands' fp32 features + the labels as bit rows (16 B per window; finetune() packs the 0/1
        # label matrix once, on first sight, during the untimed warm-up passes)
        h2d = sum(2 * sizes[c] * D * 4 + sizes[c] * ((NCLASS + 31) // 32) * 4 for c in mine)
        d2h = sum(sizes[c] * NCLASS * 4 for c in mine) + 4 * len(mine)
        for _ in range(2):
            ft.finetune(None, model, feats_host, None, optimizer, 0, None, opt_ns, "train")
        barrier()
        optimizer.grad_scale = 1.0
        t0 = time.perf_counter()
        for _ in range(steps):
            ft.finetune(None, model, feats_host, None, optimizer, 0, None, opt_ns, "train")
        barrier()
        dt = torch.tensor([(time.perf_counter() - t0) / steps], dtype=torch.float64, device=dev)
        bytes_t = torch.tensor([h2d, d2h], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            dist.all_reduce(bytes_t)
        e2e = {"value": total_edges / float(dt.item()) / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(bytes_t[0].item()),
               "d2h_bytes_per_step": int(bytes_t[1].item()), "ms_per_step": float(dt.item()) * 1e3,
               "note": "finetune() drop-in, pinned host features (fp32) and bit-packed 0/1 labels copied H2D every pass, "
                       "predictions copied D2H every pass, one optimiser step per chromosome per rank"}

    # ---- roofline of the dominant kernel: the SpMM, one launch per local chromosome, timed alone with CUDA events
    roofline = None
    if rank == 0:
        try:
            peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
            if os.path.exists(peaks_path):
                peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
            else:
                peak, peak_src = 6650.0, "B200_PROFILING.md fallback 6.65 TB/s (of fallback)"
            W = 2 * D
            outs = {c: torch.empty_like(panels[c]) for c in mine}
            for _ in range(3):
                for c in mine:
                    ops.spmm(graphs[c], panels[c].view(graphs[c].n, W), True, out=outs[c].view(graphs[c].n, W))
            torch.cuda.synchronize(dev)
            # One event pair around a back-to-back pass over the chromosomes (no host round trip between launches, so the
            # host's launch latency is not part of the interval); every launch reads a different 50+ MB panel, 1.2 GB in
            # total per pass, so nothing survives in L2 from one launch to the next.
            reps, tot_ms, tot_bytes, n_launch = 5, 0.0, 0.0, 0
            for _ in range(reps):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for c in mine:
                    ops.spmm(graphs[c], panels[c].view(graphs[c].n, W), True, out=outs[c].view(graphs[c].n, W))
                b.record()
                b.synchronize()
                tot_ms += a.elapsed_time(b)
                for c in mine:
                    g = graphs[c]
                    tot_bytes += g.nnz * (4 + 4 * W) + 4 * (g.n + 1) + 4 * W * g.n
                    n_launch += 1
            achieved = tot_bytes / (tot_ms * 1e-3) / 1e9
            # DRAM bytes per launch of the same kernel on the same workload, from the committed ncu capture
            traffic, traffic_src = None, None
            tpath = os.path.join(ROOT, "profiles", "r01b_spmm_traffic.json")
            if args.workload == "wg" and world == 1 and os.path.exists(tpath):
                tj = json.load(open(tpath))
                traffic, traffic_src = tj["dram_bytes_per_launch"], "profiles/r01b_spmm_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum, mean of %d launches)" % tj["launches"]
            roofline = {"bound": "hbm", "kernel": "spmm_pattern_kernel<2> (forward mean aggregation, width 256)",
                        "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                        "traffic_source": traffic_src,
                        "peak_source": peak_src, "avg_launch_us": tot_ms * 1e3 / n_launch,
                        "algorithmic_bytes_per_launch": tot_bytes / n_launch,
                        "note": "algorithmic bytes nnz*(4+4W)+4(N+1)+4WN; the launches of one pass run back to back between one "
                                "CUDA-event pair; Hi-C locality keeps most gathers in L1/L2, so achieved can exceed the DRAM "
                                "copy peak; see profiles/ for dram__bytes"}

        except Exception as exc:      # rank-0-only, no collective inside: a failure here must not cost the main line
            import traceback
            traceback.print_exc()
            roofline = {"bound": "hbm", "error": "%s: %s" % (type(exc).__name__, exc)}
    cpu_baseline = None
    hostbind.unbind(binding)       # the CPU baseline gets every host core back (and the JSON line no CPU list)
    if rank == 0 and not args.no_cpu_baseline:
        v, dt_cpu, e_cpu, threads = cpu_reference_run(["chr22"], steps=10, warmup=2)
        cpu_baseline = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                        "sample": "10 train steps on the chr22-sized graph (N=20000, %d stored entries), %.3f s/step; "
                                  "oracle port of the reference PyTorch CPU path incl. host process_graph" % (e_cpu, dt_cpu)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(args.workload), "d_model": D, "gcn_layers": LAYERS, "nclass": NCLASS,
                           "gate": True, "adj_type": "hic", "hicnorm": "SQRTVC", "hicsize": HIC_EDGES, "optim": "sgd",
                           "gcn_dropout": DROPOUT, "strands": 2, "total_stored_entries": total_edges,
                           "parallelism": "chromosome-sharded x%d, %d lock-step rounds per pass (balanced packing, gradient "
                                          "accumulation inside a rank's cell), one flat-gradient allreduce + optimiser step per "
                                          "round" % (world, len(schedule)),
                           "l2": "inputs larger than L2 (126 MB): %.2f GB of resident feature panels + targets cycled per "
                                 "step, ~50 panel-sized passes per chromosome" % (
                                     sum(sizes[c] for c in chroms) * (2 * D * 4 + NCLASS * 4) / 1e9),
                           "gemm_impl": args.gemm_impl, "final_loss_sum": final_loss, "host_binding_rank0": binding},
                "e2e": e2e, "gpu_launches": int(lt.item()), "clocks": clocks, "roofline": roofline,
                "cpu_baseline": cpu_baseline}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
