"""Drop-in `GraphConvolution` and `ChromeGCN` (reference: models/SubLayers.py:7-57,
models/ChromeModels.py:21-52) running on libchromegcn's sm_100a kernels.

Same constructor arguments, parameter names / shapes (state_dicts interchange with the
reference), initialisation and `forward(x_in, adj, deg, src_dict=None, return_gate=False)`
returning `(x_in, out, (g, g2), None)`.  `adj` may be the `HiCGraph` this package's
`process_graph` returns, or the torch sparse COO tensor the reference's `process_graph` returns
(converted once on the GPU and cached).  There is no eager / CPU fallback: inputs must be CUDA
fp32 and the shared library must be built.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib, ops
from .graph import HiCGraph

PARAM_ORDER = ["GC1.weight", "GC1.bias", "W1.weight", "W1.bias", "GC2.weight", "GC2.bias", "W2.weight", "W2.bias",
               "batch_norm.weight", "batch_norm.bias", "out.weight", "out.bias"]
MAX_LAYERS = _lib.MAX_LAYERS


def param_order(layers: int) -> List[str]:
    """State-dict names in the reference's order for an `layers`-layer model (GC3/W3... only in extended mode)."""
    names: List[str] = []
    for l in range(1, layers + 1):
        names += ["GC%d.weight" % l, "GC%d.bias" % l, "W%d.weight" % l, "W%d.bias" % l]
    return names + ["batch_norm.weight", "batch_norm.bias", "out.weight", "out.bias"]


def padded_classes(nclass: int) -> int:
    """Row pitch (floats) of the logit buffers: a multiple of 4 keeps rows 16-byte aligned, which puts the
    head contractions on the tcgen05 path."""
    return (nclass + 3) // 4 * 4


def bn_momentum(bn: nn.BatchNorm1d) -> float:
    """The exponential factor of the running statistics.  `momentum=None` (cumulative moving average with factor
    1 / num_batches_tracked) is not what the reference builds (models/ChromeModels.py:30: `nn.BatchNorm1d(nfeat)`)
    and is not implemented by the kernels: refuse it instead of silently using 0.1."""
    if bn.momentum is None:
        raise NotImplementedError("BatchNorm1d(momentum=None) (cumulative moving average) is not supported by the "
                                  "CUDA path; the reference uses the default momentum 0.1")
    return float(bn.momentum)


def _as_graph(adj, cache: Dict) -> HiCGraph:
    if isinstance(adj, HiCGraph):
        return adj
    if isinstance(adj, torch.Tensor) and adj.layout == torch.sparse_coo:
        key = (adj._values().data_ptr(), adj._indices().data_ptr(), adj._nnz(), tuple(adj.shape))
        g = cache.get(key)
        if g is None:
            g = HiCGraph.from_torch_coo(adj)
            if len(cache) > 64:
                cache.clear()
            cache[key] = g
        return g
    raise TypeError("adj must be a HiCGraph or a torch tensor (sparse COO or dense)")


class _GraphConvFn(torch.autograd.Function):
    """y = (A_hat x) W + b   (models/SubLayers.py:42-52), A_hat = D^-1 bin(A+I) given as a pattern."""

    @staticmethod
    def forward(ctx, x, weight, bias, graph):
        x = ops._f32c(x)
        ax = ops.spmm(graph, x, mean=True) if graph is not None else x
        y = ops.gemm_rowpanel(ax, weight.detach(), False, bias.detach() if bias is not None else None)
        ctx.graph = graph
        ctx.has_bias = bias is not None
        ctx.save_for_backward(ax, weight)
        return y

    @staticmethod
    def backward(ctx, dy):
        ax, weight = ctx.saved_tensors
        dy = ops._f32c(dy)
        graph = ctx.graph
        dw = ops.gemm_gram(ax, dy) if ctx.needs_input_grad[1] else None
        db = dy.sum(0) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        dx = None
        if ctx.needs_input_grad[0]:
            if graph is None:
                dx = ops.gemm_rowpanel(dy, weight.detach(), True)
            else:
                t = ops.gemm_rowpanel(dy, weight.detach(), True, rowscale_graph=graph, rowscale_group=1)
                dx = ops.spmm(graph, t, mean=False)
        return dx, dw, db, None


class GraphConvolution(nn.Module):
    """models/SubLayers.py:7-57.  Standalone use runs SpMM + GEMM kernels; inside `ChromeGCN` the
    whole model runs as one fused sequence instead."""

    def __init__(self, in_features, out_features, bias=True, init="xavier"):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight = nn.Parameter(torch.empty(in_features, out_features))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_features))
        else:
            self.register_parameter("bias", None)
        if init == "xavier":                      # models/SubLayers.py:32-35
            nn.init.xavier_normal_(self.weight.data, gain=0.02)
        elif init == "kaiming":                   # models/SubLayers.py:37-40
            nn.init.kaiming_normal_(self.weight.data, a=0, mode="fan_in")
        elif init == "uniform":                   # models/SubLayers.py:26-30
            stdv = 1.0 / (self.weight.size(1) ** 0.5)
            self.weight.data.uniform_(-stdv, stdv)
            if self.bias is not None:
                self.bias.data.uniform_(-stdv, stdv)
        else:
            raise NotImplementedError
        if self.bias is not None and init != "uniform":
            nn.init.constant_(self.bias.data, 0.0)
        self._graph_cache: Dict = {}

    def forward(self, input, adj, deg=None):
        _lib.require_cuda()
        graph = _as_graph(adj, self._graph_cache) if adj is not None else None
        return _GraphConvFn.apply(input, self.weight, self.bias, graph)

    def __repr__(self):
        return "%s (%d -> %d)" % (self.__class__.__name__, self.in_features, self.out_features)


def fill_params(ps: _lib.Params, tensors: Dict[str, Optional[torch.Tensor]], layers: int) -> None:
    def p(name):
        t = tensors.get(name)
        return None if t is None else _lib.ptr(t)
    for l in range(layers):
        ps.gc_w[l] = p("GC%d.weight" % (l + 1))
        ps.gc_b[l] = p("GC%d.bias" % (l + 1))
        ps.gate_w[l] = p("W%d.weight" % (l + 1))
        ps.gate_b[l] = p("W%d.bias" % (l + 1))
    ps.bn_w, ps.bn_b = p("batch_norm.weight"), p("batch_norm.bias")
    ps.out_w, ps.out_b = p("out.weight"), p("out.bias")


def build_model_struct(graph: HiCGraph, d: int, nclass: int, layers: int, strands: int, training: bool, dropout_p: float,
                       seed: int, step: int, params: Dict[str, torch.Tensor], grads: Optional[Dict[str, torch.Tensor]],
                       running_mean: torch.Tensor, running_var: torch.Tensor, num_batches: Optional[torch.Tensor],
                       x_in: torch.Tensor, x_in_grad: Optional[torch.Tensor], out: torch.Tensor,
                       gates: List[Optional[torch.Tensor]], out_grad: Optional[torch.Tensor], workspace: torch.Tensor,
                       gemm_impl: int = 0, bn_momentum: float = 0.1, bn_eps: float = 1e-5, out_ld: int = 0,
                       gate_off: bool = False) -> _lib.Model:
    m = _lib.Model()
    m.out_ld = out_ld
    m.gate_off = 1 if gate_off else 0
    m.graph = graph.c_struct()
    m.d, m.nclass, m.layers, m.strands = d, nclass, layers, strands
    m.training = 1 if training else 0
    m.gemm_impl = gemm_impl
    m.need_input_grad = 1 if x_in_grad is not None else 0
    m.dropout_p, m.bn_momentum, m.bn_eps = float(dropout_p), float(bn_momentum), float(bn_eps)
    m.seed, m.step = seed & 0xFFFFFFFFFFFFFFFF, step & 0xFFFFFFFFFFFFFFFF
    fill_params(m.params, params, layers)
    if grads is not None:
        fill_params(m.grads, grads, layers)
    m.bn_running_mean, m.bn_running_var = _lib.ptr(running_mean), _lib.ptr(running_var)
    m.bn_num_batches_tracked = _lib.ptr(num_batches)
    m.x_in, m.x_in_grad, m.out = _lib.ptr(x_in), _lib.ptr(x_in_grad), _lib.ptr(out)
    for l in range(layers):
        m.gate[l] = _lib.ptr(gates[l])
    m.out_grad = _lib.ptr(out_grad)
    m.workspace, m.workspace_bytes = workspace.data_ptr(), workspace.numel() * workspace.element_size()
    m.stream = _lib.current_stream()
    return m


class _ChromeGCNFn(torch.autograd.Function):
    """One `ChromeGCN.forward` call (strands = 1) as cgcn_model_forward / cgcn_model_backward."""

    @staticmethod
    def forward(ctx, x_in, module, graph, *param_tensors):
        lib = _lib.load()
        names = module._param_names()
        params = {k: ops._f32c(t.detach()) for k, t in zip(names, param_tensors)}
        x = ops._f32c(x_in.detach())
        n, d = x.shape
        nclass, layers = module.out.out_features, module.num_layers
        dev = x.device
        training = module.training
        with torch.cuda.device(dev):
            ws_bytes = lib.cgcn_model_workspace_bytes(n, d, nclass, layers, 1)
            if ws_bytes == 0:
                raise _lib.ChromeGCNNativeError("cgcn_model_workspace_bytes rejected n=%d d=%d" % (n, d))
            ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=dev)
            ld = padded_classes(nclass)
            out_buf = torch.empty(n, ld, dtype=torch.float32, device=dev)
            out = out_buf
            gates = [torch.empty(n, 1, dtype=torch.float32, device=dev) for _ in range(layers)]
            seed, step = module._next_dropout_counter() if training else (0, 0)
            bn = module.batch_norm
            m = build_model_struct(graph, d, nclass, layers, 1, training, module.dropout, seed, step, params, None,
                                   bn.running_mean, bn.running_var, bn.num_batches_tracked, x, None, out, gates, None, ws,
                                   module.gemm_impl, bn_momentum(bn), bn.eps, ld,
                                   module.gate_off)
            _lib.check(lib.cgcn_model_forward(C.byref(m)), "cgcn_model_forward")
        ctx.module, ctx.graph, ctx.names = module, graph, names
        ctx.cfg = (n, d, nclass, layers, training, seed, step)
        ctx.save_for_backward(x, ws, out_buf, *gates, *[params[k] for k in names])
        ctx.mark_non_differentiable(*gates)
        return (out_buf[:, :nclass], *gates)

    @staticmethod
    def backward(ctx, dout, *unused):
        lib = _lib.load()
        module, graph, names = ctx.module, ctx.graph, ctx.names
        n, d, nclass, layers, training, seed, step = ctx.cfg
        saved = ctx.saved_tensors
        x, ws, out = saved[0], saved[1], saved[2]
        gates = list(saved[3:3 + layers])
        params = dict(zip(names, saved[3 + layers:]))
        dev = x.device
        ld = out.shape[1]
        dpad = torch.zeros(n, ld, dtype=torch.float32, device=dev)
        dpad[:, :nclass] = dout
        dout = dpad
        with torch.cuda.device(dev):
            grads = {k: torch.empty_like(v) for k, v in params.items()}
            dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
            bn = module.batch_norm
            m = build_model_struct(graph, d, nclass, layers, 1, training, module.dropout, seed, step, params, grads,
                                   bn.running_mean, bn.running_var, None, x, dx, out, gates, dout, ws, module.gemm_impl,
                                   bn_momentum(bn), bn.eps, ld, module.gate_off)
            _lib.check(lib.cgcn_model_backward(C.byref(m)), "cgcn_model_backward")
        return (dx, None, None, *[grads[k] for k in names])


class ChromeGCN(nn.Module):
    """models/ChromeModels.py:21-52.  `gate` is accepted and ignored and `layers != 2` gives one
    layer, exactly like the reference.

    `extended=True` (not in the reference; the variant sweep of BASELINE.json needs "gcn_layers 3" and "gate off",
    which the reference cannot express) honours both arguments: `layers` in 1..4 builds GC1..GC{layers} / W1..W{layers}
    with the same per-layer recipe (dropout between consecutive layers), and `gate=False` makes every layer
    `x <- tanh(A_hat x W + b)` (g == 1; the W{l} parameters stay in the state_dict and receive zero gradients).
    `forward` still returns `(x_in, out, (g, g2), None)`; further gates are in `self.last_gates`."""

    def __init__(self, nfeat, nhid, nclass, dropout, gate, layers, extended: bool = False):
        super().__init__()
        if nfeat != nhid:
            raise ValueError("ChromeGCN needs nhid == nfeat (W1 = Linear(nfeat, 1) is applied to an nhid-wide tensor, "
                             "models/ChromeModels.py:24-25)")
        self.extended = bool(extended)
        self.gate_off = self.extended and not bool(gate)
        if self.extended:
            if not 1 <= int(layers) <= MAX_LAYERS:
                raise ValueError("extended ChromeGCN supports 1..%d layers" % MAX_LAYERS)
            n_layers = int(layers)
        else:
            n_layers = 2 if layers == 2 else 1
        self.GC1 = GraphConvolution(nfeat, nhid, bias=True, init="xavier")
        self.W1 = nn.Linear(nfeat, 1)
        for l in range(2, n_layers + 1):
            setattr(self, "GC%d" % l, GraphConvolution(nhid, nfeat, bias=True, init="xavier"))
            setattr(self, "W%d" % l, nn.Linear(nfeat, 1))
        self.last_gates: List[torch.Tensor] = []
        self.dropout = dropout
        self.batch_norm = nn.BatchNorm1d(nfeat)
        self.out = nn.Linear(nfeat, nclass)
        self.gemm_impl = ops.GEMM_AUTO
        self._graph_cache: Dict = {}
        self._drop_seed: Optional[int] = None
        self._drop_step = 0

    # -- helpers shared with the fused trainer
    @property
    def num_layers(self) -> int:
        n = 1
        while n < MAX_LAYERS and hasattr(self, "GC%d" % (n + 1)):
            n += 1
        return n

    def _param_names(self) -> List[str]:
        return param_order(self.num_layers)

    def _param_tensors(self) -> List[torch.Tensor]:
        sd = dict(self.named_parameters())
        return [sd[k] for k in self._param_names()]

    def _next_dropout_counter(self) -> Tuple[int, int]:
        if self._drop_seed is None:
            self._drop_seed = int(torch.randint(0, 2 ** 62, (1,)).item())     # follows torch.manual_seed
        self._drop_step += 1
        return self._drop_seed, self._drop_step

    def resolve_graph(self, adj) -> Optional[HiCGraph]:
        """The pattern graph behind `adj`, or None when `adj` needs the generic path (generic_adj.py)."""
        if isinstance(adj, torch.Tensor) and (adj.layout == torch.strided or adj.requires_grad):
            return None
        try:
            return _as_graph(adj, self._graph_cache)
        except NotImplementedError:
            if isinstance(adj, torch.Tensor):
                return None
            raise

    def forward(self, x_in, adj, deg=None, src_dict=None, return_gate=False):
        _lib.require_cuda()
        if not (isinstance(x_in, torch.Tensor) and x_in.is_cuda):
            raise _lib.ChromeGCNNativeError("ChromeGCN.forward needs CUDA inputs (no CPU fallback)")
        graph = self.resolve_graph(adj)
        if graph is None:
            # dense `adj`, `adj` that requires grad, or a sparse tensor that is not a symmetric mean-aggregation pattern
            # (scripts/visualize.py:30-45,103-111): the generic weighted-CSR path with d loss / d adj
            from .generic_adj import forward_generic
            out, gates = forward_generic(self, x_in, adj)
            self.last_gates = list(gates)
            return x_in, out, (gates[0], gates[1] if self.num_layers >= 2 else None), None
        res = _ChromeGCNFn.apply(x_in, self, graph, *self._param_tensors())
        out, g = res[0], res[1]
        g2 = res[2] if self.num_layers >= 2 else None
        self.last_gates = list(res[1:])
        return x_in, out, (g, g2), None
