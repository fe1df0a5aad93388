"""Device-resident graph handles and the drop-in `process_graph`.

Reference: `utils/util_methods.py:146-180` builds, on the host with scipy, `D^-1 bin(A + I)` as a
torch sparse COO tensor for every chromosome in every epoch and copies it to the GPU
(`finetune.py:36`).  Here the same matrix is represented by its pattern only -- an int32 CSR of
`bin(A + I)` kept on the device -- because every stored value of row i is `1/deg_i` and the
kernels derive that from `rowptr` (bit-exact).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np
import torch

from . import _lib


class HiCGraph:
    """CSR pattern of `bin(A + I)` on the GPU; what `process_graph` returns in this package.

    It stands where the reference passes the sparse `adj` tensor: `ChromeGCN.forward(x, adj, deg)`
    accepts it directly, `.cuda()` / `.to()` are no-ops, and `to_sparse_coo()` materialises the
    exact tensor the reference's `process_graph` would have produced.
    """

    def __init__(self, rowptr: torch.Tensor, colidx: torch.Tensor, n: int, nnz: int, name: str = "",
                 vals: Optional[torch.Tensor] = None, row_inv: Optional[torch.Tensor] = None):
        assert rowptr.dtype == torch.int32 and colidx.dtype == torch.int32 and rowptr.is_cuda and colidx.is_cuda
        assert (vals is None) == (row_inv is None)
        self.rowptr, self.colidx, self.n, self.nnz, self.name = rowptr, colidx, int(n), int(nnz), name
        self.vals, self.row_inv = vals, row_inv          # weighted graphs only (adj_type 'both')

    # -- torch-tensor look-alikes used by the reference's call sites
    def cuda(self, *a, **k):
        return self

    def to(self, *a, **k):
        return self

    @property
    def shape(self):
        return torch.Size((self.n, self.n))

    def size(self, dim: Optional[int] = None):
        return self.shape if dim is None else self.shape[dim]

    @property
    def device(self):
        return self.rowptr.device

    @property
    def is_cuda(self):
        return True

    def c_struct(self) -> _lib.Graph:
        return _lib.Graph(self.n, self.nnz, self.rowptr.data_ptr(), self.colidx.data_ptr(),
                          None if self.vals is None else self.vals.data_ptr(),
                          None if self.row_inv is None else self.row_inv.data_ptr())

    def degrees(self) -> torch.Tensor:
        return (self.rowptr[1:] - self.rowptr[:-1])

    def to_sparse_coo(self) -> torch.Tensor:
        """The tensor `process_graph('hic', ...)` returns in the reference: int64 indices in
        row-major / ascending-column order, fp32 values `1/deg_i`, not coalesced."""
        deg = self.degrees().to(torch.int64)
        rows = torch.repeat_interleave(torch.arange(self.n, device=self.device), deg)
        if self.vals is not None:
            vals = (self.vals[: self.nnz].double() * self.row_inv.double()[rows]).float()
        else:
            vals = (1.0 / deg.to(torch.float32))[rows]      # correctly rounded == float32(1/float64(deg))
        idx = torch.stack([rows, self.colidx[: self.nnz].to(torch.int64)])
        return torch.sparse_coo_tensor(idx, vals, (self.n, self.n), check_invariants=False)

    def csr_numpy(self):
        return self.rowptr.cpu().numpy(), self.colidx[: self.nnz].cpu().numpy()

    # -- constructors
    @classmethod
    def from_csr_pattern(cls, indptr, indices, device=None, add_selfloops: bool = True, name: str = "") -> "HiCGraph":
        """From a host CSR pattern (column-sorted).  `add_selfloops` merges the diagonal on the GPU
        (`split_adj + sparse.eye`, utils/util_methods.py:162)."""
        lib = _lib.load()
        dev = _lib.require_cuda(device)
        indptr = np.ascontiguousarray(indptr, dtype=np.int32)
        indices = np.ascontiguousarray(indices, dtype=np.int32)
        n = indptr.shape[0] - 1
        nnz = int(indptr[-1])
        rp = torch.from_numpy(indptr).to(dev)
        ci = torch.from_numpy(indices[:nnz]).to(dev) if nnz > 0 else torch.zeros(1, dtype=torch.int32, device=dev)
        if not add_selfloops:
            return cls(rp, ci, n, nnz, name)
        need = C.c_size_t(0)
        _lib.check(lib.cgcn_adj_add_selfloops_workspace_bytes(n, C.byref(need)), "cgcn_adj_add_selfloops_workspace_bytes")
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        rp_out = torch.empty(n + 1, dtype=torch.int32, device=dev)
        ci_out = torch.empty(nnz + n, dtype=torch.int32, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.cgcn_adj_add_selfloops(rp.data_ptr(), ci.data_ptr(), n, rp_out.data_ptr(), ci_out.data_ptr(),
                                                  ws.data_ptr(), need.value, _lib.current_stream()),
                       "cgcn_adj_add_selfloops")
            total = int(rp_out[-1].item())
        return cls(rp_out, ci_out, n, total, name)

    @classmethod
    def from_scipy(cls, mat, device=None, name: str = "") -> "HiCGraph":
        """`process_graph('hic')` semantics for a scipy matrix: + I, positive entries -> 1, negative -> 0
        (utils/util_methods.py:162-165); only the pattern of the positive entries survives."""
        csr = mat.tocsr()
        csr.sum_duplicates()
        if csr.nnz and not np.all(csr.data > 0):
            # entries <= 0: `+ I` can only lift a diagonal one above 0; everything else binarises to 0
            diag_fix = csr.diagonal() + 1.0 > 0
            csr = csr.multiply(csr > 0).tocsr()
            if not np.all(diag_fix):
                raise NotImplementedError("a diagonal entry <= -1 cancels the self loop; not a ChromeGCN graph")
            csr.eliminate_zeros()
        csr.sort_indices()
        return cls.from_csr_pattern(csr.indptr, csr.indices, device, True, name)

    @classmethod
    def from_torch_coo(cls, adj: torch.Tensor, name: str = "") -> "HiCGraph":
        """From the sparse COO tensor the reference's `process_graph` returns (already normalised,
        self loops included).  Verifies on the GPU that every value is `1/deg(row)` and that the
        pattern is symmetric -- the two facts the pattern-only kernels rely on."""
        lib = _lib.load()
        if not adj.is_cuda:
            adj = adj.to(_lib.require_cuda())
        dev = adj.device
        idx = adj._indices().contiguous()
        vals = adj._values().to(torch.float32).contiguous()
        nnz, n = int(vals.shape[0]), int(adj.shape[0])
        need = C.c_size_t(0)
        _lib.check(lib.cgcn_coo_to_pattern_workspace_bytes(nnz, n, C.byref(need)), "cgcn_coo_to_pattern_workspace_bytes")
        ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
        rp = torch.empty(n + 1, dtype=torch.int32, device=dev)
        ci = torch.empty(max(nnz, 1), dtype=torch.int32, device=dev)
        flags = C.c_int32(0)
        rows, cols = idx[0].contiguous(), idx[1].contiguous()
        with torch.cuda.device(dev):
            _lib.check(lib.cgcn_coo_to_pattern(rows.data_ptr(), cols.data_ptr(), vals.data_ptr(), nnz, n, rp.data_ptr(),
                                               ci.data_ptr(), C.byref(flags), ws.data_ptr(), need.value,
                                               _lib.current_stream()), "cgcn_coo_to_pattern")
        if not flags.value & 1:
            raise NotImplementedError(
                "adjacency values are not 1/deg(row): a normalised weighted tensor does not reveal the raw symmetric "
                "weights the backward pass needs; build the graph with process_graph('both', ...) or "
                "graph.weighted_from_scipy(A_raw) instead")
        if not flags.value & 2:
            raise NotImplementedError("adjacency pattern is not symmetric; the backward SpMM relies on P = P^T")
        return cls(rp, ci, n, nnz, name)


def weighted_from_scipy(mat, device=None, name: str = "") -> HiCGraph:
    """A weighted, SYMMETRIC, un-normalised scipy matrix (diagonal included) -> device graph that
    aggregates with `D^-1 A` (`normalize`, utils/util_methods.py:99-106): raw weights + 1/rowsum."""
    dev = _lib.require_cuda(device)
    csr = mat.tocsr().astype(np.float64)
    csr.sum_duplicates()
    csr.sort_indices()
    if (abs(csr - csr.T) > 0).nnz != 0:
        raise NotImplementedError("weighted adjacency must be symmetric before normalisation (backward uses A = A^T)")
    rowsum = np.asarray(csr.sum(1)).ravel()
    with np.errstate(divide="ignore"):
        r_inv = np.power(rowsum, -1.0)
    r_inv[np.isinf(r_inv)] = 0.0                               # utils/util_methods.py:103
    n = csr.shape[0]
    return HiCGraph(torch.from_numpy(csr.indptr.astype(np.int32)).to(dev), torch.from_numpy(csr.indices.astype(np.int32)).to(dev),
                    n, int(csr.nnz), name, torch.from_numpy(csr.data.astype(np.float32)).to(dev),
                    torch.from_numpy(r_inv.astype(np.float32)).to(dev))


def constant_band_pattern(constant_range: int, n: int):
    """Pattern of `create_constant_graph` (utils/util_methods.py:137-144): |i-j| <= range, i != j."""
    offs = np.arange(-constant_range, constant_range + 1)
    offs = offs[offs != 0]
    rows = np.repeat(np.arange(n), offs.shape[0])
    cols = rows + np.tile(offs, n)
    ok = (cols >= 0) & (cols < n)
    rows, cols = rows[ok], cols[ok]
    indptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(indptr, rows + 1, 1)
    return np.cumsum(indptr).astype(np.int32), cols.astype(np.int32)


def process_graph(adj_type, split_adj_dict_chrom, x_size, chrom, device=None) -> HiCGraph:
    """Drop-in for `utils.util_methods.process_graph` (utils/util_methods.py:146-180).

    Same arguments; returns a `HiCGraph` (device CSR pattern of the row-normalised matrix) instead
    of a host-built torch sparse tensor.  `hic`, `constant` and `none` are mean aggregations over a
    symmetric pattern; `both` (hic + band, not binarised: values vary inside a
    row) runs on the weighted variant of the same kernels."""
    if adj_type == "hic":
        return HiCGraph.from_scipy(split_adj_dict_chrom[chrom], device, name=str(chrom))
    if adj_type == "constant":
        ip, ix = constant_band_pattern(7, int(x_size))
        return HiCGraph.from_csr_pattern(ip, ix, device, True, name="constant")
    if adj_type == "none":
        n = int(x_size)
        return HiCGraph.from_csr_pattern(np.arange(n + 1, dtype=np.int32), np.arange(n, dtype=np.int32), device, False,
                                         name="none")
    if adj_type == "both":
        # hic + band + I, NOT binarised (utils/util_methods.py:166-169): overlapping entries weigh 2
        from scipy import sparse
        n = int(x_size)
        ip, ix = constant_band_pattern(7, n)
        band = sparse.csr_matrix((np.ones(ix.shape[0]), ix, ip), shape=(n, n))
        mat = split_adj_dict_chrom[chrom].tocsr() + band + sparse.eye(n, format="csr")
        return weighted_from_scipy(mat, device, name=str(chrom) + "+band")
    raise ValueError("unknown adj_type %r" % (adj_type,))
