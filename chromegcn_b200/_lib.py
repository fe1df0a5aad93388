"""ctypes binding of libchromegcn.so (include/chromegcn.h).

There is no CPU or PyTorch fallback behind this module: if the shared library is missing or a
call fails, a `ChromeGCNNativeError` is raised.  Importing it does not touch the GPU.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG, "lib", "libchromegcn.so")
ABI_VERSION = 8
MAX_LAYERS = 4
MAX_PEERS = 8


class ChromeGCNNativeError(RuntimeError):
    pass


class Graph(C.Structure):
    _fields_ = [("n", C.c_int32), ("nnz", C.c_int32), ("rowptr", C.c_void_p), ("colidx", C.c_void_p),
                ("vals", C.c_void_p), ("row_inv", C.c_void_p)]


class Params(C.Structure):
    _fields_ = [("gc_w", C.c_void_p * MAX_LAYERS), ("gc_b", C.c_void_p * MAX_LAYERS), ("gate_w", C.c_void_p * MAX_LAYERS),
                ("gate_b", C.c_void_p * MAX_LAYERS), ("bn_w", C.c_void_p), ("bn_b", C.c_void_p),
                ("out_w", C.c_void_p), ("out_b", C.c_void_p)]


class Model(C.Structure):
    _fields_ = [("graph", Graph),
                ("d", C.c_int32), ("nclass", C.c_int32), ("layers", C.c_int32), ("strands", C.c_int32),
                ("training", C.c_int32), ("gemm_impl", C.c_int32), ("need_input_grad", C.c_int32),
                ("out_ld", C.c_int32),
                ("dropout_p", C.c_float), ("bn_momentum", C.c_float), ("bn_eps", C.c_float), ("row_begin", C.c_int32),
                ("gate_off", C.c_int32), ("reserved0", C.c_int32),
                ("seed", C.c_uint64), ("step", C.c_uint64),
                ("params", Params), ("grads", Params),
                ("bn_running_mean", C.c_void_p), ("bn_running_var", C.c_void_p), ("bn_num_batches_tracked", C.c_void_p),
                ("x_in", C.c_void_p), ("x_in_grad", C.c_void_p), ("out", C.c_void_p), ("gate", C.c_void_p * MAX_LAYERS),
                ("out_grad", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
                ("stream", C.c_void_p),
                ("n_total", C.c_int64), ("x_full", C.c_void_p), ("bn_sums", C.c_void_p), ("peer", C.c_void_p)]


class PeerPanel(C.Structure):
    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("row_begin", C.c_int32 * (MAX_PEERS + 1)),
                ("base", C.c_void_p * MAX_PEERS)]


_P = C.c_void_p
_I32, _I64, _F32, _U64, _SZ = C.c_int32, C.c_int64, C.c_float, C.c_uint64, C.c_size_t

# name -> (restype, argtypes); every symbol include/chromegcn.h declares
PROTOTYPES = {
    "cgcn_abi_version": (C.c_int, []),
    "cgcn_last_error": (C.c_char_p, []),
    "cgcn_device_info": (C.c_int, [C.POINTER(_I32)] * 3),
    "cgcn_sizeof": (_SZ, [_I32]),
    "cgcn_launch_count": (_I64, []),
    "cgcn_adj_build_workspace_bytes": (C.c_int, [_I64, _I64, _I64, C.POINTER(_SZ)]),
    "cgcn_adj_build": (C.c_int, [_P, _P, _P, _I64, _P, _I64, _P, _I64, _I64, _I64, _I32, _P, _P, _I64,
                                 C.POINTER(_I64), _P, _SZ, _P]),
    "cgcn_adj_add_selfloops_workspace_bytes": (C.c_int, [_I32, C.POINTER(_SZ)]),
    "cgcn_adj_add_selfloops": (C.c_int, [_P, _P, _I32, _P, _P, _P, _SZ, _P]),
    "cgcn_coo_to_pattern_workspace_bytes": (C.c_int, [_I64, _I64, C.POINTER(_SZ)]),
    "cgcn_coo_to_pattern": (C.c_int, [_P, _P, _P, _I64, _I32, _P, _P, C.POINTER(_I32), _P, _SZ, _P]),
    "cgcn_spmm": (C.c_int, [C.POINTER(Graph), _P, _P, _I32, _I32, _P, _P]),
    "cgcn_sddmm": (C.c_int, [C.POINTER(Graph), _P, _P, _I32, _P, _P]),
    "cgcn_spmm_peer": (C.c_int, [C.POINTER(Graph), C.POINTER(PeerPanel), _P, _I32, _I32, _P, _P]),
    "cgcn_peer_alloc": (C.c_int, [_SZ, C.POINTER(_P), C.c_char_p]),
    "cgcn_peer_open": (C.c_int, [C.c_char_p, C.POINTER(_P)]),
    "cgcn_peer_close": (C.c_int, [_P]),
    "cgcn_peer_free": (C.c_int, [_P]),
    "cgcn_peer_publish": (C.c_int, [_P, _P, _SZ, _P]),
    "cgcn_gcn_layer_fwd": (C.c_int, [C.POINTER(Graph), _I32, _P, _P, _P, _P, _P, _P, _I32, _F32, _U64, _U64, _I32, _P, _P, _P, _P,
                                     _P, C.POINTER(_I32), _P]),
    "cgcn_gcn_layer_bwd": (C.c_int, [C.POINTER(Graph), _I32, _P, _P, _P, _P, _P, _P, _P, _I32, _F32, _U64, _U64, _I32, _P, _P, _P,
                                     _P, C.POINTER(_I32), _P]),
    "cgcn_gemm_rowpanel": (C.c_int, [_P, _I64, _P, _I32, _P, _P, _I64, _I64, _I32, _I32, _P, _P, _I32, _I32, _P, _SZ, _P]),
    "cgcn_gemm_gram_workspace_bytes": (_SZ, [_I64]),
    "cgcn_gemm_gram": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _I64, _I32, _I32, _I32, _I32, _P, _SZ, _P]),
    "cgcn_model_workspace_bytes": (_SZ, [_I32, _I32, _I32, _I32, _I32]),
    "cgcn_model_forward": (C.c_int, [C.POINTER(Model)]),
    "cgcn_model_backward": (C.c_int, [C.POINTER(Model)]),
    "cgcn_bce_workspace_bytes": (_SZ, [_I32, _I32]),
    "cgcn_bce_loss": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I64, _P, _P, _P, _P, _SZ, _P]),
    "cgcn_model_phase": (C.c_int, [C.POINTER(Model), _I32, _I32, C.POINTER(_P)]),
    "cgcn_bce_loss_bits": (C.c_int, [_P, _P, _I32, _I32, _I32, _I32, _I64, _P, _P, _P, _P, _SZ, _P]),
    "cgcn_text_count_rows": (C.c_int, [C.c_char_p, _I32, _P]),
    "cgcn_contacts_parse": (C.c_int, [C.c_char_p, _I64, _P, _P, _P, _P, _I32]),
    "cgcn_vector_parse": (C.c_int, [C.c_char_p, _I64, _P, _P, _I32]),
    "cgcn_bed_starts_parse": (C.c_int, [C.c_char_p, C.c_char_p, _I64, _P, _P, _P, _I32]),
    "cgcn_label_metrics_workspace_bytes": (_SZ, [_I64, _I32]),
    "cgcn_label_metrics": (C.c_int, [_P, _I64, _P, _I64, _P, _I64, _I32, C.c_double, _P, _P, _SZ, _P]),
    "cgcn_train_step": (C.c_int, [C.POINTER(Model), _P, _P, _P, _P]),
    "cgcn_train_step_bits": (C.c_int, [C.POINTER(Model), _P, _P, _P, _P]),
    "cgcn_sgd_step": (C.c_int, [_P, _P, _P, _I64, _F32, _F32, _F32, _F32, _P]),
    "cgcn_adam_step": (C.c_int, [_P, _P, _P, _P, _I64, _F32, _F32, _F32, _F32, _I64, _F32, _P]),
    "cgcn_comm_unique_id": (C.c_int, [C.c_char_p]),
    "cgcn_comm_init": (C.c_int, [C.POINTER(_P), C.c_char_p, _I32, _I32]),
    "cgcn_comm_destroy": (C.c_int, [_P]),
    "cgcn_comm_allreduce_sum": (C.c_int, [_P, _P, _SZ, _P]),
    "cgcn_comm_allgather": (C.c_int, [_P, _P, _P, _SZ, _P]),
    "cgcn_membw_read": (C.c_int, [_P, _SZ, _I32, _P, _P]),
    "cgcn_interleave_strands": (C.c_int, [C.POINTER(_P), _I32, _I32, _I32, _P, _P]),
    "cgcn_deinterleave_strands": (C.c_int, [_P, _I32, _I32, _I32, C.POINTER(_P), _P]),
    "cgcn_dropout_mask": (C.c_int, [_P, _I32, _I32, _I32, _F32, _U64, _U64, _I32, _P]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load the shared library (once) and bind every prototype.  No GPU needed."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ChromeGCNNativeError(
            "libchromegcn.so is not built (%s). Build it with `python -m chromegcn_b200.build`; "
            "there is no CPU fallback for the ChromeGCN path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)          # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.cgcn_abi_version() != ABI_VERSION:
        raise ChromeGCNNativeError("libchromegcn.so ABI %d != binding ABI %d" % (lib.cgcn_abi_version(), ABI_VERSION))
    for which, st in ((0, Graph), (1, Params), (2, Model), (3, PeerPanel)):
        if lib.cgcn_sizeof(which) != C.sizeof(st):
            raise ChromeGCNNativeError("struct %s: C sizeof %d != ctypes sizeof %d" %
                                       (st.__name__, lib.cgcn_sizeof(which), C.sizeof(st)))
    _lib = lib
    return lib


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = load().cgcn_last_error()
        raise ChromeGCNNativeError("%s failed (status %d): %s" % (what or "libchromegcn call", status,
                                                                  msg.decode("utf-8", "replace") if msg else ""))


def launch_count() -> int:
    return int(load().cgcn_launch_count())


def require_cuda(device=None):
    """The product path has no CPU implementation: fail loudly when there is no GPU."""
    import torch
    if not torch.cuda.is_available():
        raise ChromeGCNNativeError("chromegcn_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)


def ptr(t) -> Optional[int]:
    """Device pointer of a contiguous torch tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_contiguous():
        raise ChromeGCNNativeError("non-contiguous tensor passed to libchromegcn")
    return t.data_ptr()


def current_stream() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream
