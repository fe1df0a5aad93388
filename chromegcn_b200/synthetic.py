"""Synthetic GM12878-shaped inputs for the ChromeGCN chromosome-model path.

There is no network and no ENCODE / Juicer data on the benchmark box, so every
measured number is taken on inputs produced here:

* windows   -- a sorted random subset of 1 kb bin starts (multiples of 1000, the
               shape `data/1create_windows.py:49-59` of the reference produces),
               N_c = round(20000 * len_c / len_chr22) windows per chromosome;
* contacts  -- a Juicer `RAWobserved`-like triplet list `(bin1, bin2, count)` in
               ascending `(bin1, bin2)` order, genomic distance log-uniform (contact
               density P(s) ~ 1/s), integer counts that decay with distance (the input
               of `data/7create_graph_new.py:67-91`);
* norm      -- a Juicer `*.SQRTVCnorm`-like per-bin vector with a few NaN and
               0.0 entries (`data/7create_graph_new.py:51-65`);
* features  -- N(0,1) fp32 `[N, 128]` forward / reverse-complement features and
               Bernoulli(0.05) targets, the `chrom_feature_dict` layout of
               `utils/util_methods.py:183-199`.

Everything is numpy / torch-CPU and deterministic in the chromosome index.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence

import numpy as np

# hg19 chromosome lengths (bp); chr1..chr22 + chrX.
HG19_LENGTHS: Dict[str, int] = {
    "chr1": 249250621, "chr2": 243199373, "chr3": 198022430, "chr4": 191154276,
    "chr5": 180915260, "chr6": 171115067, "chr7": 159138663, "chr8": 146364022,
    "chr9": 141213431, "chr10": 135534747, "chr11": 135006516, "chr12": 133851895,
    "chr13": 115169878, "chr14": 107349540, "chr15": 102531392, "chr16": 90354753,
    "chr17": 81195210, "chr18": 78077248, "chr19": 59128983, "chr20": 63025520,
    "chr21": 48129895, "chr22": 51304566, "chrX": 155270560,
}
AUTOSOMES: List[str] = ["chr%d" % i for i in range(1, 23)]
WHOLE_GENOME: List[str] = AUTOSOMES + ["chrX"]
# reference split (data/create_data.py:40-45)
VALID_CHROMS = ["chr3", "chr12", "chr17"]
TEST_CHROMS = ["chr1", "chr8", "chr21"]

BIN_BP = 1000            # GM12878 Hi-C resolution is 1 kb (data/create_data.py:52-55)
CHR22_WINDOWS = 20000    # config 1 of BASELINE.json
NCLASS = 103
D_MODEL = 128


def chrom_index(chrom: str) -> int:
    return 23 if chrom == "chrX" else int(chrom[3:])


def num_windows(chrom: str, scale: float = 1.0) -> int:
    n = round(CHR22_WINDOWS * HG19_LENGTHS[chrom] / HG19_LENGTHS["chr22"] * scale)
    return max(int(n), 8)


@dataclass
class SyntheticHiC:
    """One chromosome's raw inputs to the adjacency build."""
    chrom: str
    window_starts: np.ndarray   # int64 [N], sorted, unique, multiples of 1000
    bin1: np.ndarray            # int64 [M]
    bin2: np.ndarray            # int64 [M]
    val: np.ndarray             # float64 [M]
    norm: np.ndarray            # float64 [n_bins] (may hold NaN / 0.0)
    resolution_kb: int = 1


def make_windows(chrom: str, scale: float = 1.0, n_windows: Optional[int] = None,
                 n_bins: Optional[int] = None) -> np.ndarray:
    c = chrom_index(chrom)
    rng = np.random.default_rng(1000 + c)
    if n_bins is None:
        n_bins = int(HG19_LENGTHS[chrom] * scale) // BIN_BP
    n = num_windows(chrom, scale) if n_windows is None else n_windows
    n = min(n, n_bins)
    picks = rng.choice(n_bins, size=n, replace=False)
    picks.sort()
    return picks.astype(np.int64) * BIN_BP


def make_hic(chrom: str, hic_edges: int = 500000, scale: float = 1.0,
             candidates_per_edge: float = 8.0, n_windows: Optional[int] = None,
             n_bins: Optional[int] = None, max_dist_bins: int = 2000) -> SyntheticHiC:
    """Distance-decay contact list for one chromosome.

    `candidates_per_edge * K` draws (K = hic_edges / 2 undirected pairs): bin1 uniform over the
    window bins, genomic distance log-uniform in [1, max_dist_bins] bins (contact density
    P(s) ~ 1/s), about 15 % of the rows then get a non-window bin1 so the window filter of
    `data/7create_graph_new.py:78` has something to reject; integer counts
    max(1, Poisson(60/s) + Geometric(0.3)) so near-diagonal contacts dominate the top-K but ties are
    everywhere; rows ordered by (bin1, bin2) with duplicates summed, like a Juicer dump.
    """
    c = chrom_index(chrom)
    if n_bins is None:
        n_bins = int(HG19_LENGTHS[chrom] * scale) // BIN_BP
    windows = make_windows(chrom, scale, n_windows, n_bins)
    wbins = windows // BIN_BP
    rng = np.random.default_rng(2000 + c)
    k_pairs = int(hic_edges / 2.0)
    m = int(candidates_per_edge * k_pairs)
    b1 = wbins[rng.integers(0, wbins.shape[0], size=m)]
    off = rng.random(m) < 0.15
    b1 = np.where(off, rng.integers(0, n_bins - 1, size=m, dtype=np.int64), b1)
    dist = np.floor(np.exp(rng.random(m) * np.log(max_dist_bins + 1.0))).astype(np.int64)
    dist = np.clip(dist, 1, max_dist_bins)
    b2 = b1 + dist
    keep = b2 < n_bins
    b1, b2, dist = b1[keep], b2[keep], dist[keep]
    counts = np.maximum(1, rng.poisson(60.0 / dist) + rng.geometric(0.3, size=b1.shape[0]) - 1).astype(np.float64)
    # Juicer dumps are ordered by (bin1, bin2); duplicates collapse (counts add)
    key = b1 * np.int64(n_bins) + b2
    order = np.argsort(key, kind="stable")
    key, counts = key[order], counts[order]
    uniq, first = np.unique(key, return_index=True)
    summed = np.add.reduceat(counts, first)
    b1 = (uniq // n_bins) * BIN_BP
    b2 = (uniq % n_bins) * BIN_BP
    norm = rng.lognormal(0.0, 0.3, size=n_bins)
    u = rng.random(n_bins)
    norm[u < 0.01] = np.nan
    norm[(u >= 0.01) & (u < 0.011)] = 0.0
    return SyntheticHiC(chrom, windows, b1.astype(np.int64), b2.astype(np.int64),
                        summed.astype(np.float64), norm.astype(np.float64), 1)


def make_features(chrom: str, n: int, d: int = D_MODEL, nclass: int = NCLASS):
    """`chrom_feature_dict[chrom]` entry: {'forward','backward','target'} torch CPU tensors."""
    import torch
    c = chrom_index(chrom)
    g = torch.Generator().manual_seed(3000 + c)
    x_f = torch.randn(n, d, generator=g, dtype=torch.float32)
    x_r = torch.randn(n, d, generator=g, dtype=torch.float32)
    tgt = (torch.rand(n, nclass, generator=g) < 0.05).to(torch.float32)
    return {"forward": x_f, "backward": x_r, "target": tgt}


def make_pattern_direct(n: int, k_pairs: int, seed: int, max_dist: int = 2000):
    """Large symmetric pattern straight in window-index space (stress graph, N = 1e6):
    the same distance law as `make_hic` without the 8x candidate list.  Returns the
    scipy CSR (float64 ones, zero diagonal) the reference's pickles would hold."""
    from scipy import sparse
    rng = np.random.default_rng(seed)
    m = int(k_pairs * 1.5)
    i = rng.integers(0, n - 1, size=m, dtype=np.int64)
    dist = np.clip(np.floor(np.exp(rng.random(m) * np.log(max_dist + 1.0))).astype(np.int64), 1, max_dist)
    j = i + dist
    keep = j < n
    key = np.unique(i[keep] * np.int64(n) + j[keep])
    if key.shape[0] > k_pairs:
        key = rng.choice(key, size=k_pairs, replace=False)
    i, j = key // n, key % n
    rows = np.concatenate([i, j])
    cols = np.concatenate([j, i])
    a = sparse.csr_matrix((np.ones(rows.shape[0]), (rows, cols)), shape=(n, n))
    a.sum_duplicates()
    a.data[:] = 1.0
    a.sort_indices()
    return a


def write_juicer_files(root: str, cell_type: str, hics: Sequence[SyntheticHiC], norm_name: str = "SQRTVC",
                       bed_name: str = "chipseq_windows.bed", write_sorted: bool = False) -> Dict[str, str]:
    """Lay the synthetic inputs out on disk the way `create_graph` expects them
    (`data/7create_graph_new.py:140-147,173-179`): a bed file of windows under
    `<root>/out/` and per-chromosome `RAWobserved` / `*norm` text files."""
    import os
    out_root = os.path.join(root, "out")
    os.makedirs(os.path.join(out_root, "hic"), exist_ok=True)
    hic_root = os.path.join(root, "hic")
    with open(os.path.join(out_root, bed_name), "w") as fp:
        for h in hics:
            for s in h.window_starts.tolist():
                fp.write("%s\t%d\t%d\tA0\t0\t.\t0\t0\t0\t0\n" % (h.chrom, s, s + BIN_BP))
    for h in hics:
        res = str(h.resolution_kb)
        d = os.path.join(hic_root, cell_type + "_combined", res + "kb_resolution_intrachromosomal",
                         h.chrom, "MAPQGE30")
        os.makedirs(d, exist_ok=True)
        raw = os.path.join(d, "%s_%skb.RAWobserved" % (h.chrom, res))
        with open(raw, "w") as fp:
            for a, b, v in zip(h.bin1.tolist(), h.bin2.tolist(), h.val.tolist()):
                fp.write("%d\t%d\t%r\n" % (a, b, v))
        if write_sorted:
            # GNU `sort -r -k3 -n` equivalent is emulated by the caller; here: stable desc by value
            order = np.argsort(-h.val, kind="stable")
            with open(raw + ".sorted", "w") as fp:
                for t in order.tolist():
                    fp.write("%d\t%d\t%r\n" % (int(h.bin1[t]), int(h.bin2[t]), float(h.val[t])))
        with open(os.path.join(d, "%s_%skb.%snorm" % (h.chrom, res, norm_name)), "w") as fp:
            for v in h.norm.tolist():
                fp.write(("NaN" if v != v else repr(v)) + "\n")
    return {"output_root": out_root, "hic_root": hic_root}
