"""Drop-in for the offline adjacency build `data/7create_graph_new.py`.

Same entry point `create_graph(args)` and the same `args` fields (`output_root, use_all_windows,
hic_root, cell_type, resolution, hic_edges, norm, chroms, valid_chroms, test_chroms`), same input
files (Juicer `RAWobserved` / `*norm` dumps, the windows bed file) and the same outputs: three
pickles `{split}_graphs_{hic_edges}_{norm}norm.pkl` of `{chrom: scipy.sparse.csr_matrix float64}`
(`data/7create_graph_new.py:147-149,197-202`), byte-identical `indptr` / `indices`.

What changes is where the work happens: the text files are parsed once into arrays by the library's
multi-threaded memory-mapped parser (`cgcn_contacts_parse` / `cgcn_vector_parse`, csrc/ingest.cu; every value
equals Python's `int(str)` / `float(str)`; `CGCN_TEXT_PARSER=pandas` selects the pandas C parser with
`float_precision='round_trip'` instead), and filter,
fp64 normalisation, dict-semantics dedup, stable top-K, symmetrisation and CSR assembly run on the
GPU (`cgcn_adj_build`) instead of a per-row Python loop over a dict, a Python sort and a dense
N x N matrix (`:67-120`).
"""
from __future__ import annotations

import os
import pickle
from typing import Dict

import numpy as np


def read_window_starts(bed_file: str, chroms) -> Dict[str, np.ndarray]:
    """`create_bin_dict` (data/7create_graph_new.py:14-47): per chromosome the sorted unique start
    positions of the bed rows; a window's index is its rank."""
    if _native_parser():
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        names = list(chroms)
        joined = "\n".join(names).encode()
        rows = C.c_int64(0)
        rc = lib.cgcn_bed_starts_parse(os.fsencode(bed_file), joined, 0, None, None, C.byref(rows), 0)
        if rc not in (0, -5):                                  # -5 = CGCN_ERR_CAPACITY: the sizing call
            _lib.check(rc, "cgcn_bed_starts_parse")
        n = int(rows.value)
        ci, st = np.empty(n, dtype=np.int32), np.empty(n, dtype=np.int64)
        if n:
            _lib.check(lib.cgcn_bed_starts_parse(os.fsencode(bed_file), joined, n, ci.ctypes.data, st.ctypes.data, C.byref(rows), 0),
                       "cgcn_bed_starts_parse")
        return {c: np.unique(st[ci == i]) for i, c in enumerate(names)}
    import pandas as pd
    df = pd.read_csv(bed_file, sep="\t", header=None, usecols=[0, 1], names=["chrom", "start"],
                     dtype={"chrom": str, "start": np.int64})
    out = {}
    wanted = set(chroms)
    for chrom, grp in df[df["chrom"].isin(wanted)].groupby("chrom"):
        out[chrom] = np.unique(grp["start"].to_numpy(dtype=np.int64))
    for c in chroms:
        out.setdefault(c, np.zeros(0, dtype=np.int64))
    return out


def _native_parser() -> bool:
    return os.environ.get("CGCN_TEXT_PARSER", "native") != "pandas"


def _count_rows(lib, path: str, threads: int) -> int:
    import ctypes as C
    from . import _lib
    rows = C.c_int64(0)
    _lib.check(lib.cgcn_text_count_rows(os.fsencode(path), threads, C.byref(rows)), "cgcn_text_count_rows")
    return int(rows.value)


def read_norm_vector(path: str, threads: int = 0) -> np.ndarray:
    """`get_normalization_values` (:51-65) minus the NaN / 0 -> inf substitution, which the kernel applies."""
    if _native_parser():
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        n = _count_rows(lib, path, threads)
        out = np.empty(n, dtype=np.float64)
        rows = C.c_int64(0)
        _lib.check(lib.cgcn_vector_parse(os.fsencode(path), n, out.ctypes.data, C.byref(rows), threads), "cgcn_vector_parse")
        return out[: int(rows.value)]
    import pandas as pd
    return pd.read_csv(path, sep="\t", header=None, usecols=[0], names=["v"], dtype={"v": np.float64},
                       float_precision="round_trip", na_values=["NaN", "nan"], keep_default_na=True)["v"].to_numpy()


def read_contacts(path: str, threads: int = 0):
    """The `start_pos1 \\t start_pos2 \\t val` triplets of a RAWobserved dump (:71-76)."""
    if _native_parser():
        import ctypes as C
        from . import _lib
        lib = _lib.load()
        n = _count_rows(lib, path, threads)
        b1, b2, v = np.empty(n, dtype=np.int64), np.empty(n, dtype=np.int64), np.empty(n, dtype=np.float64)
        rows = C.c_int64(0)
        _lib.check(lib.cgcn_contacts_parse(os.fsencode(path), n, b1.ctypes.data, b2.ctypes.data, v.ctypes.data, C.byref(rows),
                                           threads), "cgcn_contacts_parse")
        k = int(rows.value)
        return b1[:k], b2[:k], v[:k]
    import pandas as pd
    df = pd.read_csv(path, sep="\t", header=None, usecols=[0, 1, 2], names=["b1", "b2", "v"],
                     dtype={"b1": np.int64, "b2": np.int64, "v": np.float64}, float_precision="round_trip")
    return df["b1"].to_numpy(), df["b2"].to_numpy(), df["v"].to_numpy()


def build_chromosome(window_starts, bin1, bin2, val, norm, resolution, hic_edges, device=None):
    """One chromosome: arrays in, `scipy.sparse.csr_matrix` (float64 ones) out (:182-188)."""
    from scipy import sparse
    from . import ops
    indptr, indices = ops.adjacency_build(window_starts, bin1, bin2, val, norm, int(resolution), int(hic_edges), device)
    n = indptr.shape[0] - 1
    return sparse.csr_matrix((np.ones(indices.shape[0], dtype=np.float64), indices, indptr), shape=(n, n))


def create_graph(args, device=None):
    output_root = args.output_root
    bed = os.path.join(output_root, "windows.bed" if args.use_all_windows else "chipseq_windows.bed")
    hic_root = os.path.join(args.hic_root, args.cell_type + "_combined", str(args.resolution) + "kb_resolution_intrachromosomal/")
    names = {s: os.path.join(output_root, "hic/%s_graphs_%s_%snorm.pkl" % (s, str(args.hic_edges), args.norm))
             for s in ("train", "valid", "test")}
    print("\nInputs\n| " + bed + "\n| " + hic_root + "\n\nOutputs")
    for s in ("train", "valid", "test"):
        print("| " + names[s])
    starts = read_window_starts(bed, args.chroms)
    dicts = {"train": {}, "valid": {}, "test": {}}
    for chrom in args.chroms:
        print(chrom)
        base = os.path.join(hic_root, chrom + "/MAPQGE30/" + chrom + "_" + str(args.resolution) + "kb.")
        if args.norm != "":
            norm = read_norm_vector(base + args.norm + "norm")
            b1, b2, v = read_contacts(base + "RAWobserved")
        else:                                   # pre-sorted dump, first K accepted rows (:177-179, :88-89)
            norm = None
            b1, b2, v = read_contacts(base + "RAWobserved.sorted")
        adj = build_chromosome(starts[chrom], b1, b2, v, norm, args.resolution, args.hic_edges, device)
        split = "test" if chrom in args.test_chroms else ("valid" if chrom in args.valid_chroms else "train")
        dicts[split][chrom] = adj
    os.makedirs(os.path.join(output_root, "hic"), exist_ok=True)
    for s in ("train", "valid", "test"):
        with open(names[s], "wb") as fp:
            pickle.dump(dicts[s], fp)
    return dicts
