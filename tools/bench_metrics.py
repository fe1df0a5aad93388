"""Whole-genome split metrics: cgcn_label_metrics on the GPU vs the reference's sklearn route (utils/metrics.py) on the
host, same [sum N, nclass] probability / label matrices.  The sklearn leg runs on a bounded sample of labels."""
import json, sys, time, warnings
import numpy as np, torch
sys.path.insert(0, ".")
from chromegcn_b200 import metrics, ops

n, c, sample = 1183638, 103, 6
rng = np.random.default_rng(0)
t = (rng.random((n, c)) < 0.05).astype(np.float32)
p = (1 / (1 + np.exp(-(2.0 * t + rng.standard_normal((n, c)).astype(np.float32) - 2.5)))).astype(np.float32)
dev = torch.device("cuda", 0)
pd, bits = torch.from_numpy(p).to(dev), ops.pack_targets(torch.from_numpy(t)).to(dev)
for _ in range(2):
    got = metrics.label_metrics_device(pd, bits)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    got = metrics.label_metrics_device(pd, bits)
torch.cuda.synchronize()
gpu_s = (time.perf_counter() - t0) / 5
with warnings.catch_warnings():
    warnings.simplefilter("ignore")
    t0 = time.perf_counter()
    a = metrics.auroc(t[:, :sample], p[:, :sample])[3]
    b = metrics.aupr(t[:, :sample], p[:, :sample])[3]
    f = metrics.fdr(t[:, :sample], p[:, :sample])[3]
    cpu_s = (time.perf_counter() - t0) * c / sample
err = max(np.abs(got["auroc"][:sample] - a).max(), np.abs(got["aupr"][:sample] - b).max(), np.abs(got["fdr"][:sample] - f).max())
print(json.dumps({"n": n, "nclass": c, "gpu_ms": gpu_s * 1e3, "sklearn_s_extrapolated_from_%d_labels" % sample: cpu_s,
                  "speedup": cpu_s / gpu_s, "max_abs_diff_on_sample": float(err),
                  "key_sort_bytes": n * c * 8}))
