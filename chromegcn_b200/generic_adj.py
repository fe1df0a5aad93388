"""Arbitrary adjacency matrices: dense tensors, asymmetric or re-weighted sparse tensors, and d loss / d adj.

The training path hands `ChromeGCN.forward` the row-normalised `D^-1 bin(A + I)` of `process_graph` (a symmetric
PATTERN: graph.HiCGraph, the fused kernels).  The reference's analysis code also calls the model with
  * a DENSE `adj` that requires grad and reads `adj.grad` (A-saliency, scripts/visualize.py:30-45:
    `adj_grad = |adj * adj.grad|`, so only the gradient on the non-zero support matters), and
  * re-normalised sparse tensors with masked-out entries -- arbitrary values, asymmetric pattern
    (scripts/visualize.py:103-111).
Neither fits the pattern-only kernels.  This module runs them on the library's weighted kernels: `cgcn_spmm` on a CSR
with values (forward  A S,  backward  A^T G  through the transposed CSR) and `cgcn_sddmm` for the value gradient
`d a_ij = <G_i, S_j>`, with the dense contractions on `cgcn_gemm_rowpanel` / `cgcn_gemm_gram`; the element-wise rest of
the model (tanh, gate, ReLU, BatchNorm, dropout) is torch under autograd -- this is the analysis path, one model call
at a time, not the training loop.  Same math and operation order as models/SubLayers.py:42-52 / models/ChromeModels.py:
34-52 (`A (x W) + b`).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch
import torch.nn.functional as F

from . import _lib, ops
from .graph import HiCGraph


class GenericAdj:
    """CSR (int32) of an arbitrary square matrix and of its transpose, on the device.  `entry` maps the CSR order back
    to the caller's storage: flat index `i * n + j` of a dense tensor, or the position in a sparse tensor's value
    array; `tperm` maps the transposed CSR's entries to positions of the forward CSR."""

    def __init__(self, n: int, rows: torch.Tensor, cols: torch.Tensor, entry: torch.Tensor):
        dev = rows.device
        self.n, self.nnz, self.entry = n, int(rows.shape[0]), entry

        def csr(r, c):
            key = r.to(torch.int64) * n + c.to(torch.int64)
            order = torch.argsort(key, stable=True)
            counts = torch.bincount(r[order].to(torch.int64), minlength=n)
            rowptr = torch.zeros(n + 1, dtype=torch.int64, device=dev)
            rowptr[1:] = torch.cumsum(counts, 0)
            return rowptr.to(torch.int32).contiguous(), c[order].to(torch.int32).contiguous(), order

        self.rowptr, self.colidx, order = csr(rows, cols)
        self.entry = entry[order]
        r_sorted, c_sorted = rows[order], cols[order]
        self.rowptr_t, self.colidx_t, self.tperm = csr(c_sorted, r_sorted)
        self.ones = torch.ones(n, dtype=torch.float32, device=dev)
        if self.nnz == 0:
            self.colidx = torch.zeros(1, dtype=torch.int32, device=dev)
            self.colidx_t = torch.zeros(1, dtype=torch.int32, device=dev)

    def graph(self, vals: torch.Tensor, transposed: bool = False) -> HiCGraph:
        if transposed:
            return HiCGraph(self.rowptr_t, self.colidx_t, self.n, self.nnz, "adj^T", vals[self.tperm].contiguous(), self.ones)
        return HiCGraph(self.rowptr, self.colidx, self.n, self.nnz, "adj", vals.contiguous(), self.ones)


def from_tensor(adj: torch.Tensor):
    """`(GenericAdj, values in CSR order as a differentiable function of adj)` for a dense or sparse square tensor."""
    dev = _lib.require_cuda(adj.device if adj.is_cuda else None)
    n = int(adj.shape[0])
    if adj.layout == torch.strided:
        a = adj.to(dev) if not adj.is_cuda else adj
        idx = torch.nonzero(a.detach(), as_tuple=False)
        rows, cols = idx[:, 0].contiguous(), idx[:, 1].contiguous()
        ga = GenericAdj(n, rows, cols, rows.to(torch.int64) * n + cols.to(torch.int64))
        vals = a.reshape(-1)[ga.entry].to(torch.float32)          # differentiable gather: adj.grad lands on the support
        return ga, vals
    a = adj.to(dev) if not adj.is_cuda else adj
    a = a.coalesce() if not a.is_coalesced() else a
    idx = a.indices()
    ga = GenericAdj(n, idx[0].contiguous(), idx[1].contiguous(), torch.arange(idx.shape[1], device=dev))
    return ga, a.values()[ga.entry].to(torch.float32)


class _SpmmValuesFn(torch.autograd.Function):
    """Y = A S for a CSR with values; gradients for S (A^T G) and for the values (SDDMM)."""

    @staticmethod
    def forward(ctx, vals, s, ga: GenericAdj):
        s = ops._f32c(s)
        vals = ops._f32c(vals.detach())
        y = ops.spmm(ga.graph(vals), s, mean=False)
        ctx.ga = ga
        ctx.save_for_backward(vals, s)
        return y

    @staticmethod
    def backward(ctx, g):
        vals, s = ctx.saved_tensors
        ga = ctx.ga
        g = ops._f32c(g)
        ds = ops.spmm(ga.graph(vals, transposed=True), g, mean=False) if ctx.needs_input_grad[1] else None
        dv = None
        if ctx.needs_input_grad[0]:
            lib = _lib.load()
            dv = torch.empty(max(ga.nnz, 1), dtype=torch.float32, device=g.device)[: ga.nnz]
            gs = ga.graph(vals).c_struct()
            with torch.cuda.device(g.device):
                _lib.check(lib.cgcn_sddmm(C.byref(gs), g.data_ptr(), s.data_ptr(), int(s.shape[1]), dv.data_ptr(),
                                          _lib.current_stream()), "cgcn_sddmm")
        return dv, ds, None


def forward_generic(model, x_in: torch.Tensor, adj: torch.Tensor):
    """`ChromeGCN.forward` (models/ChromeModels.py:34-52) for an arbitrary `adj` tensor; returns `(out, gates)`."""
    from .chrome_models import _GraphConvFn
    ga, vals = from_tensor(adj)
    x = ops._f32c(x_in)
    gates = []
    for l in range(1, model.num_layers + 1):
        gc, wl = getattr(model, "GC%d" % l), getattr(model, "W%d" % l)
        if l > 1:
            x = F.dropout(x, model.dropout, training=model.training)                 # :42
        support = _GraphConvFn.apply(x, gc.weight, None, None)                        # SubLayers.py:43  x W
        z = torch.tanh(_SpmmValuesFn.apply(vals, support, ga) + gc.bias)             # :46,50 ; ChromeModels.py:38
        if model.gate_off:
            g = torch.ones(z.shape[0], 1, dtype=z.dtype, device=z.device)
            x = z
        else:
            g = torch.sigmoid(wl(z))                                                  # :39
            x = (1 - g) * x + g * z                                                   # :40
        gates.append(g)
    x = model.batch_norm(F.relu(x))                                                   # :48-49
    x = F.dropout(x, model.dropout, training=model.training)                          # :50
    return model.out(x), gates                                                        # :51
