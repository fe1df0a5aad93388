"""Per-kernel share and algorithmic bandwidth of one whole-genome train step, from an ncu launch list
(`ncu --metrics gpu__time_duration.sum ... bench.py --steps 1`).  Algorithmic bytes are analytic (DESIGN.md
section 4) for the WG workload: 23 chromosomes, sum N = 1 183 638, sum nnz(A_hat) = 12 683 638, S = 2, d = 128,
nclass 103 (logit row pitch 104).  ncu times are cold-cache and serialised: read the SHARES, not the absolutes."""
import csv, collections, re, sys, json

path = sys.argv[1]
peak = 6541.1
N, NNZ, S, D, C, LD = 1183638, 12683638, 2, 128, 103, 104
P = N * S * D * 4                    # one feature panel
O = N * S * LD * 4                   # one logit panel
lines = [l for l in open(path) if not l.startswith("==")]
tot = collections.defaultdict(float); cnt = collections.Counter()
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("cgcn::", "").replace("tc::", "")
    tot[name] += float(row["Metric Value"].replace(",", "")); cnt[name] += 1
spmm = 3 * (NNZ * (4 + 4 * S * D) + 4 * N + P) + P
bytes_ = {
    "spmm_pattern_kernel<2, 0>": (spmm, "3 x [nnz(4+4W) + 4N + 4WN] + residual"),
    "gemm_rowpanel_tc_kernel": (3 * 2 * P + 2 * (P + O), "y1, y2, t2: P in + P out; head, d hb: P + logits"),
    "gemm_gram_tc_kernel": (2 * 2 * P + P + O, "dW1, dW2: 2P; dWout: P + logits"),
    "gate_fwd_kernel<1, 2, 0>": (4 * P, "y, x in; z, x' out"),
    "gate_fwd_kernel<1, 2, 1>": (4 * P, "same + BN column sums"),
    "gate_bwd_kernel<1, 2, 1>": (6 * P, "d hb, h, z, x in; dy, dxd out"),
    "gate_bwd_kernel<1, 2, 0>": (4 * P, "dx, z, x in; dy out"),
    "bn_bwd_reduce_kernel<1, 2>": (2 * P, "d hb, h in"),
    "bn_apply_kernel": (2 * P, "h in, hb out"),
    "bce_kernel<2>": (2 * O + 2 * N * C * 4, "logits in, d logits out, targets in, probs out"),
    "colsum_partial_kernel": (O, "d logits in"),
}
T = sum(tot.values())
print("| kernel | launches | time (ms) | share | algorithmic GB | GB/s | of %.0f GB/s | bytes counted |" % peak)
print("|---|---|---|---|---|---|---|---|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    b = bytes_.get(k)
    if b:
        gbs = b[0] / (v * 1e-9) / 1e9
        print("| `%s` | %d | %.2f | %.1f %% | %.2f | %.0f | %.2f | %s |" % (k, cnt[k], v / 1e6, 100 * v / T, b[0] / 1e9, gbs, gbs / peak, b[1]))
    else:
        print("| `%s` | %d | %.2f | %.1f %% | - | - | - | small / latency bound |" % (k, cnt[k], v / 1e6, 100 * v / T))
print("| **total** | %d | %.2f | 100 %% | | | | |" % (sum(cnt.values()), T / 1e6))
