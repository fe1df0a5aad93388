"""Stress graph (BASELINE.json configs[3], single-GPU part): one synthetic chromosome with N = 1e6 windows and
~51 M stored entries; SpMM forward timed alone at row widths 128..1024 floats (d_model 128 / 512, one or two
strands).  Panels are 0.5-4 GB, far beyond L2: this is the HBM-streamed regime.  Prints one JSON line."""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chromegcn_b200 import ops, synthetic
from chromegcn_b200.graph import HiCGraph

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1_000_000
k_pairs = int(float(sys.argv[2])) if len(sys.argv) > 2 else 25_000_000
dev = torch.device("cuda", 0)
t0 = time.time()
a = synthetic.make_pattern_direct(n, k_pairs, seed=77)
g = HiCGraph.from_csr_pattern(a.indptr, a.indices, dev)
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6650.0
res = {"n": g.n, "nnz": g.nnz, "build_s": round(time.time() - t0, 1), "peak_gbs": peak, "runs": []}
for width in (128, 256, 512, 1024):
    x = torch.randn(g.n, width, device=dev)
    out = torch.empty_like(x)
    for _ in range(3):
        ops.spmm(g, x, True, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        ops.spmm(g, x, True, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    alg = g.nnz * (4 + 4 * width) + 4 * (g.n + 1) + 4 * width * g.n
    comp = 4 * g.nnz + 4 * (g.n + 1) + 8 * width * g.n
    res["runs"].append({"width": width, "us": round(us, 1), "alg_GBs": round(alg / us / 1e3, 1), "frac_of_peak": round(alg / us / 1e3 / peak, 3),
                        "compulsory_GBs": round(comp / us / 1e3, 1), "edges_per_s_G": round(g.nnz / us / 1e3, 2)})
    del x, out
print(json.dumps(res))
