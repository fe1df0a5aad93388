"""CPU: the C-ABI library loads, exports every symbol include/chromegcn.h declares, the ctypes
structs match the C layout, and the product path refuses to run without a GPU (no fallback)."""
import os
import re
import subprocess

import pytest
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from chromegcn_b200 import build, _lib
    build.build()
    return _lib.load()


def _declared():
    hdr = open(os.path.join(REPO, "include", "chromegcn.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(cgcn_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from chromegcn_b200 import _lib
    names = _declared()
    assert len(names) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (cgcn_[a-z0-9_]+)", out))
    assert set(names) <= exported, sorted(set(names) - exported)
    assert set(names) == set(_lib.PROTOTYPES), sorted(set(names) ^ set(_lib.PROTOTYPES))
    assert not re.search(r"torch|at::|c10::", out)            # plain C ABI, no torch types behind it


def test_abi_version_and_struct_layout(lib):
    import ctypes as C
    from chromegcn_b200 import _lib
    assert lib.cgcn_abi_version() == _lib.ABI_VERSION == 8
    assert lib.cgcn_sizeof(0) == C.sizeof(_lib.Graph)
    assert lib.cgcn_sizeof(1) == C.sizeof(_lib.Params)
    assert lib.cgcn_sizeof(2) == C.sizeof(_lib.Model)


def test_sass_is_sm100a(lib):
    from chromegcn_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(lib):
    from chromegcn_b200 import _lib, ops
    from chromegcn_b200.chrome_models import ChromeGCN
    from chromegcn_b200.graph import process_graph
    m = ChromeGCN(128, 128, 7, 0.2, True, 2)
    assert sorted(k for k, _ in m.named_parameters()) == sorted(
        ["GC1.weight", "GC1.bias", "W1.weight", "W1.bias", "GC2.weight", "GC2.bias", "W2.weight", "W2.bias",
         "batch_norm.weight", "batch_norm.bias", "out.weight", "out.bias"])
    with pytest.raises(_lib.ChromeGCNNativeError):
        m(torch.randn(4, 128), None, None)
    with pytest.raises(_lib.ChromeGCNNativeError):
        process_graph("none", None, 4, "chr1")
    with pytest.raises(_lib.ChromeGCNNativeError):
        ops.dropout_mask(4, 1, 128, 0.1, 1, 1, 0)


def test_reference_state_dict_loads():
    """Checkpoints interchange: the oracle / reference state_dict loads into the drop-in module."""
    from chromegcn_b200.chrome_models import ChromeGCN
    from oracle.gcn import ChromeGCNOracle
    for layers in (1, 2, 3):
        ref = ChromeGCNOracle(128, 128, 11, 0.2, True, layers)
        ours = ChromeGCN(128, 128, 11, 0.2, True, layers)
        res = ours.load_state_dict(ref.state_dict())
        assert not res.missing_keys and not res.unexpected_keys
        assert ours.num_layers == (2 if layers == 2 else 1)       # reference quirk: layers != 2 -> one layer
        assert [tuple(v.shape) for v in ours.state_dict().values()] == [tuple(v.shape) for v in ref.state_dict().values()]
    # reference initialisation: xavier_normal_(gain=0.02), zero bias (models/SubLayers.py:32-35)
    torch.manual_seed(0)
    m = ChromeGCN(128, 128, 11, 0.2, True, 2)
    assert float(m.GC1.bias.abs().max()) == 0.0
    assert abs(float(m.GC1.weight.std()) - 0.02 * (2.0 / 256) ** 0.5) < 2e-4


def test_cli_flags_match_reference():
    """config_args.py:4-54: the single-dash flags of the README GCN recipe parse to the reference's derived fields."""
    import argparse
    from chromegcn_b200.config_args import config_args, get_args
    argv = ("-load_pretrained -chrome_model gcn -gate -gcn_layers 2 -adj_type hic -hicnorm SQRTVC -hicsize 500000 "
            "-optim sgd -lr 0.25 -gcn_dropout 0.2 -epochs 1000 -dataroot /data -results_dir /res").split()
    opt = config_args(get_args(argparse.ArgumentParser(), argv))
    assert opt.graph_root == "/data/GM12878/1000/hic" and opt.batch_size == 512 and opt.hicsize == "500000"
    assert opt.model_name.endswith(".finetune.lr2_002.gcndrop_20.adam.gcn.layers_2.gate.adj_hic.norm_SQRTVC")
    assert opt.model_name.startswith("/res/GM12878/graph.expecto.128.bsz_64.loss_ce.sgd.lr_25.drop_10_10")


def test_pack_targets_bit_layout():
    """Host side of cgcn_*_bits: bit c of row r == label (r, c); soft labels are refused."""
    import numpy as np
    import torch
    from chromegcn_b200 import ops
    rng = np.random.default_rng(0)
    for c in (1, 31, 32, 33, 103, 128, 200):
        t = (rng.random((50, c)) < 0.4).astype(np.float32)
        b = ops.pack_targets(torch.from_numpy(t))
        assert b.dtype == torch.int32 and tuple(b.shape) == (50, (c + 31) // 32)
        w = b.numpy().view(np.uint32)
        back = ((w[:, np.arange(c) >> 5] >> (np.arange(c) & 31).astype(np.uint32)) & 1).astype(np.float32)
        assert np.array_equal(back, t)
        if c % 32:
            assert not (w[:, -1] >> np.uint32(c % 32)).any()       # padding bits stay clear
    t[0, 0] = 0.25
    assert ops.pack_targets(torch.from_numpy(t)) is None


def test_comm_entry_points_validate_arguments_without_a_gpu():
    """cgcn_comm_* (collectives for a non-Python host): argument errors are reported before NCCL is touched."""
    import ctypes as C
    from chromegcn_b200 import _lib
    lib = _lib.load()
    handle = C.c_void_p()
    ident = C.create_string_buffer(128)
    assert lib.cgcn_comm_init(C.byref(handle), ident, 2, 5) == -1
    assert b"rank 5 of world 2" in lib.cgcn_last_error()
    assert lib.cgcn_comm_init(None, ident, 1, 0) == -1
    assert lib.cgcn_comm_allreduce_sum(None, None, 4, None) == -1
    assert lib.cgcn_comm_allgather(None, None, None, 4, None) == -1
    assert lib.cgcn_comm_destroy(None) == 0


def _header_prototypes():
    """name -> list of parameter declarations, parsed from include/chromegcn.h."""
    hdr = open(os.path.join(REPO, "include", "chromegcn.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|int64_t|size_t|const char\s*\*)\s+(cgcn_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", hdr, flags=re.S):
        params = " ".join(m.group(2).split())
        protos[m.group(1)] = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
    return protos


def test_ctypes_prototypes_agree_with_the_header():
    """Every binding in _lib.PROTOTYPES has the header's parameter count, and each parameter's ctypes class matches the
    C declaration's class (pointer / 32-bit / 64-bit / float): a drifted binding would corrupt the call frame silently."""
    import ctypes as C
    from chromegcn_b200 import _lib
    protos = _header_prototypes()
    assert set(protos) == set(_lib.PROTOTYPES)

    def c_class(decl):
        if "*" in decl or "[" in decl or re.search(r"\bcgcn_(stream|comm)_t\b", decl):
            return "ptr"
        if re.search(r"\b(int64_t|uint64_t|size_t)\b", decl):
            return "i64"
        if re.search(r"\b(float)\b", decl):
            return "f32"
        if re.search(r"\b(double)\b", decl):
            return "f64"
        if re.search(r"\b(int32_t|uint32_t|int)\b", decl):
            return "i32"
        raise AssertionError("unclassified parameter: %r" % decl)

    def py_class(t):
        if t in (C.c_void_p, C.c_char_p) or isinstance(t, type(C.POINTER(C.c_int))):
            return "ptr"
        return {C.c_int64: "i64", C.c_uint64: "i64", C.c_size_t: "i64", C.c_float: "f32", C.c_double: "f64",
                C.c_int32: "i32", C.c_uint32: "i32", C.c_int: "i32"}[t]

    for name, params in sorted(protos.items()):
        _, argtypes = _lib.PROTOTYPES[name]
        assert len(argtypes) == len(params), (name, len(argtypes), params)
        for t, decl in zip(argtypes, params):
            assert py_class(t) == c_class(decl), (name, decl, t)
