"""GPU parity: each C-ABI kernel against the CPU oracle / fp64 torch on the same seeded inputs.

Tolerances (max-norm relative, `oracle.gcn.max_rel`):
  * integer work (adjacency, patterns): bit-exact;
  * SpMM: 2e-6 (same fp32 summation order as the reference's per-row accumulation);
  * fp32 FFMA contractions: 2e-6 against fp64;
  * tcgen05 3xTF32 contractions: 1e-5 against fp64 (north_star budget), and reported.
"""
import os

import numpy as np
import pytest
import torch

from oracle import adjacency as oadj
from oracle import gcn as ogcn

pytestmark = pytest.mark.gpu

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CHROMS = ["chr1", "chr2", "chr3", "chr22"]


def _dev():
    return torch.device("cuda", 0)


def _random_pattern(n, avg_deg, seed, hub=None):
    rng = np.random.default_rng(seed)
    m = n * avg_deg // 2
    i = rng.integers(0, n, m)
    j = np.clip(i + rng.integers(1, 200, m), 0, n - 1)
    if hub is not None:
        hi = np.full(hub, 7)
        hj = rng.choice(n, hub, replace=False)
        i, j = np.concatenate([i, hi]), np.concatenate([j, hj])
    keep = i != j
    ip, ix = oadj._pairs_to_csr(n, i[keep], j[keep])
    return ip, ix


def _graph(ip, ix):
    from chromegcn_b200.graph import HiCGraph
    return HiCGraph.from_csr_pattern(ip, ix, _dev(), add_selfloops=True)


@pytest.mark.parametrize("width", [128, 256, 512, 1024])
@pytest.mark.parametrize("mean", [True, False])
def test_spmm_matches_oracle(width, mean):
    from chromegcn_b200 import ops
    n = 3001
    ip, ix = _random_pattern(n, 12, 5, hub=1500)          # row 7 is a hub (> LONG_ROW): block-cooperative path
    g = _graph(ip, ix)
    rp, ci = oadj.pattern_with_selfloops(ip, ix)
    assert np.array_equal(g.csr_numpy()[0], rp) and np.array_equal(g.csr_numpy()[1], ci)
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(n, width, generator=gen)
    res = torch.randn(n, width, generator=gen)
    a = ogcn.coo_adjacency(ip, ix, torch.float64).coalesce()
    if not mean:
        a = torch.sparse_coo_tensor(a.indices(), torch.ones_like(a.values()), a.shape)
    want = torch.sparse.mm(a, x.double())
    got = ops.spmm(g, x.to(_dev()), mean=mean).cpu()
    assert ogcn.max_rel(got, want) <= 2e-6
    got2 = ops.spmm(g, x.to(_dev()), mean=mean, residual=res.to(_dev())).cpu()
    assert ogcn.max_rel(got2, want + res.double()) <= 2e-6


@pytest.mark.parametrize("width", [128, 256])
@pytest.mark.parametrize("n", [1013, 37, 4500])
def test_spmm_near_diagonal_graph_matches_oracle_and_peer_kernel(width, n):
    """Near-diagonal (Hi-C-like) graphs: short-range pairs, 5 % long-range pairs, one hub row, isolated rows, n not
    a multiple of the CTA's row count.  Checked against the fp64 oracle and, bit for bit, against the peer-memory
    kernel with one block (both kernels sum a row's neighbours in CSR order)."""
    from chromegcn_b200 import ops
    rng = np.random.default_rng(n + width)
    m = n * 8
    i = rng.integers(0, n, m)
    j = np.clip(i + rng.integers(-20, 21, m), 0, n - 1)
    far = rng.random(m) < 0.05
    j[far] = rng.integers(0, n, int(far.sum()))
    if n > 800:                                                           # row 5: hub (> LONG_ROW entries)
        hj = rng.choice(n, 700, replace=False)
        i, j = np.concatenate([i, np.full(700, 5)]), np.concatenate([j, hj])
    keep = (i != j) & (i % 97 != 3)                                       # rows 3, 100, 197, ... keep only their self loop
    ip, ix = oadj._pairs_to_csr(n, i[keep], j[keep])
    g = _graph(ip, ix)
    gen = torch.Generator().manual_seed(n)
    x = torch.randn(n, width, generator=gen)
    res = torch.randn(n, width, generator=gen)
    a = ogcn.coo_adjacency(ip, ix, torch.float64).coalesce()
    want = torch.sparse.mm(a, x.double())
    got = ops.spmm(g, x.to(_dev()), mean=True, residual=res.to(_dev()))
    assert ogcn.max_rel(got.cpu(), want + res.double()) <= 2e-6
    same_order = ops.spmm_peer(g, [x.to(_dev())], 0, mean=True, residual=res.to(_dev()))
    assert torch.equal(got, same_order)
    plain = ops.spmm(g, x.to(_dev()), mean=False)
    assert torch.equal(plain, ops.spmm_peer(g, [x.to(_dev())], 0, mean=False))


def test_spmm_isolated_rows_and_tiny_graph():
    from chromegcn_b200 import ops
    ip = np.array([0, 0, 1, 2, 2], dtype=np.int32)        # rows 0 and 3 isolated -> self loop only
    ix = np.array([2, 1], dtype=np.int32)
    g = _graph(ip, ix)
    x = torch.arange(4 * 128, dtype=torch.float32).view(4, 128)
    got = ops.spmm(g, x.to(_dev())).cpu()
    want = torch.stack([x[0], (x[1] + x[2]) / 2, (x[1] + x[2]) / 2, x[3]])
    assert torch.equal(got, want)


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("m,k,n,bt,bias,scale", [(1000, 128, 128, False, True, False), (777, 128, 103, True, True, False),
                                                  (515, 103, 128, False, False, False), (4099, 128, 128, True, False, True),
                                                  (130, 128, 128, False, False, False), (64, 37, 5, True, True, False),
                                                  # weight operands wider than one 128 x 128 block (d_model 256 / 512)
                                                  (1000, 512, 512, False, True, False), (777, 512, 103, True, True, False),
                                                  (515, 103, 512, False, False, False), (600, 256, 256, True, False, True),
                                                  (300, 200, 300, False, True, False)])
def test_gemm_rowpanel(impl, m, k, n, bt, bias, scale):
    from chromegcn_b200 import ops
    gen = torch.Generator().manual_seed(m)
    a = torch.randn(m, k, generator=gen)
    b = torch.randn(n, k, generator=gen) if bt else torch.randn(k, n, generator=gen)
    bi = torch.randn(n, generator=gen) if bias else None
    want = a.double() @ (b.double().t() if bt else b.double())
    g = None
    if scale:
        ip, ix = _random_pattern(m, 6, 3)
        g = _graph(ip, ix)
        deg = torch.from_numpy(np.diff(oadj.pattern_with_selfloops(ip, ix)[0])).double()
        want = want / deg[:, None]
    if bias:
        want = want + bi.double()
    got = ops.gemm_rowpanel(a.to(_dev()), b.to(_dev()), bt, bi.to(_dev()) if bias else None, g, 1, impl).cpu()
    tol = 2e-6 if impl == 1 else 1e-5
    err = ogcn.max_rel(got, want)
    print("rowpanel impl=%d m=%d k=%d n=%d err=%.2e" % (impl, m, k, n, err))
    assert err <= tol


@pytest.mark.parametrize("impl", [1, 0])
@pytest.mark.parametrize("m,ka,nb", [(5000, 128, 128), (70001, 103, 128), (33, 128, 128), (2049, 128, 128),
                                     (5000, 512, 512), (3000, 103, 512), (2000, 256, 128), (700, 300, 200)])
def test_gemm_gram(impl, m, ka, nb):
    from chromegcn_b200 import ops
    gen = torch.Generator().manual_seed(m)
    a = torch.randn(m, ka, generator=gen)
    b = torch.randn(m, nb, generator=gen)
    want = a.double().t() @ b.double()
    got = ops.gemm_gram(a.to(_dev()), b.to(_dev()), impl).cpu()
    tol = 2e-6 if impl == 1 else 1e-5
    # reduction over m signed terms: scale by the no-cancellation magnitude sqrt(m)
    err = float((got.double() - want).abs().max()) / float(want.abs().max())
    print("gram impl=%d m=%d err=%.2e" % (impl, m, err))
    assert err <= tol * 4


def test_bce_loss_and_gradient():
    from chromegcn_b200 import ops
    gen = torch.Generator().manual_seed(2)
    n, c, s = 1001, 103, 2
    out = torch.randn(n, s, c, generator=gen) * 3
    out[0, :, 0] = 40.0
    out[1, :, 1] = -40.0
    tgt = (torch.rand(n, c, generator=gen) < 0.2).float()
    o64 = out.double().requires_grad_(True)
    pred = o64.mean(1)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(pred, tgt.double())
    loss.backward()
    acc = torch.zeros(1, device=_dev())
    probs, grad = ops.bce_loss(out.to(_dev()), tgt.to(_dev()), s, acc)
    assert abs(acc.item() - loss.item()) <= 2e-6 * abs(loss.item())
    assert ogcn.max_rel(probs.cpu(), torch.sigmoid(pred)) <= 2e-6
    assert ogcn.max_rel(grad.cpu(), o64.grad) <= 2e-6
    ops.bce_loss(out.to(_dev()), tgt.to(_dev()), s, acc, want_probs=False, want_grad=False)   # accumulates
    assert abs(acc.item() - 2 * loss.item()) <= 4e-6 * abs(loss.item())


@pytest.mark.parametrize("c", [103, 1, 32, 33, 128, 200])
def test_bce_loss_bit_packed_labels_identical_to_float(c):
    """`cgcn_bce_loss_bits` (labels as bit rows) must reproduce `cgcn_bce_loss` bit for bit."""
    from chromegcn_b200 import ops
    gen = torch.Generator().manual_seed(20 + c)
    n, s = 777, 2
    ld = (c + 3) // 4 * 4
    out = torch.zeros(n, s, ld)
    out[:, :, :c] = torch.randn(n, s, c, generator=gen) * 3
    tgt = (torch.rand(n, c, generator=gen) < 0.3).float()
    bits = ops.pack_targets(tgt)
    assert bits is not None and bits.dtype == torch.int32 and tuple(bits.shape) == (n, (c + 31) // 32)
    a0, a1 = torch.zeros(1, device=_dev()), torch.zeros(1, device=_dev())
    p0, g0 = ops.bce_loss(out.to(_dev()), tgt.to(_dev()), s, a0)
    p1, g1 = ops.bce_loss(out.to(_dev()), bits.to(_dev()), s, a1, nclass=c)
    assert torch.equal(p0, p1) and torch.equal(g0, g1) and a0.item() == a1.item()
    soft = tgt.clone()
    soft[3, 0] = 0.5
    assert ops.pack_targets(soft) is None                      # soft labels keep the float path


@pytest.mark.parametrize("kind", ["sgd", "adam"])
def test_optimizer_kernels_match_torch(kind):
    from chromegcn_b200 import ops
    gen = torch.Generator().manual_seed(3)
    p0 = torch.randn(5000, generator=gen)
    ref = torch.nn.Parameter(p0.clone().double())
    opt = (torch.optim.SGD([ref], lr=0.25, momentum=0.9, weight_decay=1e-6) if kind == "sgd"
           else torch.optim.Adam([ref], lr=2e-3, betas=(0.9, 0.98)))
    p = p0.clone().to(_dev())
    b1 = torch.zeros_like(p)
    b2 = torch.zeros_like(p)
    for step in range(1, 6):
        g = torch.randn(5000, generator=gen)
        ref.grad = g.double()
        opt.step()
        if kind == "sgd":
            ops.sgd_step(p, g.to(_dev()), b1, 0.25)
        else:
            ops.adam_step(p, g.to(_dev()), b1, b2, 2e-3, step)
    assert ogcn.max_rel(p.cpu(), ref.data) <= 2e-6


def test_interleave_roundtrip_and_dropout_mask():
    from chromegcn_b200 import ops
    gen = torch.Generator().manual_seed(4)
    a, b = torch.randn(333, 128, generator=gen), torch.randn(333, 128, generator=gen)
    panel = ops.interleave_strands([a.to(_dev()), b.to(_dev())])
    assert torch.equal(panel.cpu(), torch.stack([a, b], 1))
    ra, rb = ops.deinterleave_strands(panel)
    assert torch.equal(ra.cpu(), a) and torch.equal(rb.cpu(), b)
    m1 = ops.dropout_mask(4000, 2, 128, 0.2, 123, 5, 0)
    m2 = ops.dropout_mask(4000, 2, 128, 0.2, 123, 5, 0)
    m3 = ops.dropout_mask(4000, 2, 128, 0.2, 123, 6, 0)
    m4 = ops.dropout_mask(4000, 2, 128, 0.2, 123, 5, 1)
    assert torch.equal(m1, m2) and not torch.equal(m1, m3) and not torch.equal(m1, m4)
    vals = torch.unique(m1).cpu().tolist()
    assert len(vals) == 2 and vals[0] == 0.0 and abs(vals[1] - 1.25) < 1e-6
    keep = float((m1 > 0).float().mean())
    assert abs(keep - 0.8) < 3e-3                       # 1.0e6 draws: sigma = 4e-4


# ------------------------------------------------------------------------------ adjacency (bit-exact)
@pytest.mark.parametrize("fname", ["adjacency_SQRTVC_1200.npz", "adjacency_none_1200.npz", "adjacency_SQRTVC_41.npz",
                                   "adjacency_SQRTVC_1000000.npz"])
def test_adjacency_build_golden(fname):
    from chromegcn_b200 import ops
    z = np.load(os.path.join(GOLDEN, fname))
    use_norm = str(z["norm_name"]) != ""
    for c in CHROMS:
        ip, ix = ops.adjacency_build(z[c + "_windows"], z[c + "_bin1"], z[c + "_bin2"], z[c + "_val"],
                                     z[c + "_norm"] if use_norm else None, 1, int(z["hic_edges"]))
        assert ip.dtype == np.int32 and ix.dtype == np.int32
        assert np.array_equal(ip, z[c + "_indptr"]), (fname, c)
        assert np.array_equal(ix, z[c + "_indices"]), (fname, c)


@pytest.mark.parametrize("chrom,hic_edges,use_norm", [("chr22", 500000, True), ("chr22", 125000, True),
                                                      ("chr20", 1000000, True), ("chr21", 250000, False)])
def test_adjacency_build_full_size_vs_oracle(chrom, hic_edges, use_norm):
    """Config-1 sized input (1.5 M contact rows, 20 k windows): byte-identical CSR + structural properties."""
    from chromegcn_b200 import ops, synthetic
    h = synthetic.make_hic(chrom, hic_edges=hic_edges)
    b1, b2, v = h.bin1, h.bin2, h.val
    if not use_norm:
        order = np.argsort(-v, kind="stable")
        b1, b2, v = b1[order], b2[order], v[order]
    norm = h.norm if use_norm else None
    ip, ix = ops.adjacency_build(h.window_starts, b1, b2, v, norm, 1, hic_edges)
    wp, wx = oadj.build_adjacency_numpy(h.window_starts, b1, b2, v, norm, 1, hic_edges)
    assert np.array_equal(ip, wp) and np.array_equal(ix, wx)
    n = ip.shape[0] - 1
    rows = np.repeat(np.arange(n), np.diff(ip))
    assert not np.any(rows == ix)
    assert ip[-1] <= hic_edges
    fwd = rows.astype(np.int64) * n + ix
    bwd = ix.astype(np.int64) * n + rows
    assert np.array_equal(np.sort(fwd), np.sort(bwd))          # symmetric
    assert np.all(np.diff(fwd) > 0)                            # sorted, unique


def test_adjacency_edge_cases():
    from chromegcn_b200 import ops, _lib
    w = np.array([0, 1000, 2000, 5000], dtype=np.int64)
    norm = np.ones(6)
    # no acceptable contact at all
    ip, ix = ops.adjacency_build(w, np.array([3000, 0]), np.array([4000, 0]), np.array([1.0, 9.0]), norm, 1, 10)
    assert np.array_equal(ip, np.zeros(5, dtype=np.int32)) and ix.shape[0] == 0
    # empty contact list
    ip, ix = ops.adjacency_build(w, np.zeros(0, np.int64), np.zeros(0, np.int64), np.zeros(0), norm, 1, 10)
    assert np.array_equal(ip, np.zeros(5, dtype=np.int32)) and ix.shape[0] == 0
    # (a,b) and (b,a) selected together collapse; K larger than the candidates keeps everything
    ip, ix = ops.adjacency_build(w, np.array([0, 1000, 2000]), np.array([1000, 0, 5000]), np.array([3.0, 2.0, 1.0]),
                                 norm, 1, 100)
    wp, wx = oadj.build_adjacency_loops(w, np.array([0, 1000, 2000]), np.array([1000, 0, 5000]),
                                        np.array([3.0, 2.0, 1.0]), norm, 1, 100)
    assert np.array_equal(ip, wp) and np.array_equal(ix, wx) and ip[-1] == 4
    # a bin beyond the norm vector is the reference's IndexError
    with pytest.raises(_lib.ChromeGCNNativeError):
        ops.adjacency_build(w, np.array([0]), np.array([5000]), np.array([1.0]), np.ones(3), 1, 10)
    # NaN contact value: undefined order in the reference -> rejected
    with pytest.raises(_lib.ChromeGCNNativeError):
        ops.adjacency_build(w, np.array([0]), np.array([5000]), np.array([np.nan]), norm, 1, 10)


def test_adjacency_build_random_small_cases_vs_loop_oracle():
    """150 seeded adversarial cases (heavy value ties, duplicate and reversed keys, NaN / 0 norm entries, diagonal and
    non-window rows, K from 0 to more than the candidates, with and without the norm vector) through cgcn_adj_build,
    byte-compared with the record-at-a-time restatement of data/7create_graph_new.py:67-120."""
    from chromegcn_b200 import ops
    rng = np.random.default_rng(2024)
    vals_pool = np.array([0.0, 1.0, 2.0, 2.0, 3.0, 5.5, 7.0])
    norm_pool = np.array([1.0, 0.5, 2.0, np.nan, 0.0, 1.25])
    for case in range(150):
        n_bins = int(rng.integers(4, 60))
        starts = np.sort(rng.choice(n_bins, int(rng.integers(1, n_bins + 1)), replace=False)).astype(np.int64) * 1000
        m = int(rng.integers(0, 200))
        b1 = rng.integers(0, n_bins, m).astype(np.int64) * 1000
        b2 = rng.integers(0, n_bins, m).astype(np.int64) * 1000
        v = vals_pool[rng.integers(0, len(vals_pool), m)]
        norm = norm_pool[rng.integers(0, len(norm_pool), n_bins)] if case % 3 else None
        hic_edges = int(rng.integers(0, 2 * m + 5))
        want = oadj.build_adjacency_loops(starts, b1, b2, v, norm, 1, hic_edges)
        got = ops.adjacency_build(starts, b1, b2, v, norm, 1, hic_edges, _dev())
        assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1]), (case, n_bins, m, hic_edges)


def test_process_graph_golden():
    from scipy import sparse
    from chromegcn_b200.graph import process_graph, HiCGraph
    z = np.load(os.path.join(GOLDEN, "process_graph.npz"))
    for c in CHROMS:
        ip, ix = z[c + "_indptr"], z[c + "_indices"]
        n = ip.shape[0] - 1
        csr = sparse.csr_matrix((np.ones(ix.shape[0]), ix, ip), shape=(n, n))
        g = process_graph("hic", {c: csr}, n, c)
        assert isinstance(g, HiCGraph) and g.cuda() is g
        rp, ci = g.csr_numpy()
        wrp, wci = oadj.pattern_with_selfloops(ip, ix)
        assert np.array_equal(rp, wrp) and np.array_equal(ci, wci)
        coo = g.to_sparse_coo()
        assert not coo.is_coalesced()
        assert np.array_equal(coo._indices()[0].cpu().numpy(), z[c + "_coo_rows"])
        assert np.array_equal(coo._indices()[1].cpu().numpy(), z[c + "_coo_cols"])
        assert np.array_equal(coo._values().cpu().numpy(), z[c + "_coo_vals"])        # bit-exact fp32 1/deg
        # and back: the reference's tensor -> pattern
        g2 = HiCGraph.from_torch_coo(coo)
        assert np.array_equal(g2.csr_numpy()[0], wrp) and np.array_equal(g2.csr_numpy()[1], wci)


def test_process_graph_other_adj_types():
    """adj_type constant / none (pattern kernels) and both (weighted kernels) against the reference's tensors."""
    from scipy import sparse
    from chromegcn_b200.graph import process_graph
    from chromegcn_b200 import ops
    z = np.load(os.path.join(GOLDEN, "adj_types.npz"))
    n = int(z["n"])
    csr = sparse.csr_matrix((np.ones(z["indices"].shape[0]), z["indices"], z["indptr"]), shape=(n, n))
    gen = torch.Generator().manual_seed(9)
    x = torch.randn(n, 256, generator=gen)
    for t in ("constant", "both", "none"):
        g = process_graph(t, {"chrZ": csr}, n, "chrZ")
        coo = g.to_sparse_coo().coalesce()
        assert np.array_equal(coo.indices()[0].cpu().numpy(), z[t + "_rows"])
        assert np.array_equal(coo.indices()[1].cpu().numpy(), z[t + "_cols"])
        assert np.allclose(coo.values().cpu().numpy(), z[t + "_vals"], rtol=2e-7, atol=0)
        assert (g.vals is not None) == (t == "both")
        ref = torch.sparse_coo_tensor(torch.from_numpy(np.vstack((z[t + "_rows"], z[t + "_cols"]))),
                                      torch.from_numpy(z[t + "_vals"]).double(), (n, n))
        want = torch.sparse.mm(ref, x.double())
        assert ogcn.max_rel(ops.spmm(g, x.to(_dev()), mean=True).cpu(), want) <= 2e-6
        # backward operator: A_hat^T G = A (row_inv .* G)
        wantT = ref.to_dense().t() @ x.double()
        if g.vals is None:
            scaled = x / g.degrees().cpu().float()[:, None]
        else:
            scaled = x * g.row_inv.cpu()[:, None]
        assert ogcn.max_rel(ops.spmm(g, scaled.to(_dev()), mean=False).cpu(), wantT) <= 2e-6


def test_model_on_weighted_graph_matches_reference():
    """ChromeGCN step with adj_type 'both' (weighted SpMM, row_inv row scaling in the backward GEMM)."""
    from scipy import sparse
    from chromegcn_b200.chrome_models import ChromeGCN
    from chromegcn_b200.engine import ChromosomeEngine
    from chromegcn_b200.graph import process_graph
    z = np.load(os.path.join(GOLDEN, "adj_types.npz"))
    n = int(z["n"])
    csr = sparse.csr_matrix((np.ones(z["indices"].shape[0]), z["indices"], z["indptr"]), shape=(n, n))
    g = process_graph("both", {"c": csr}, n, "c")
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd0.")}
    nclass = sd["out.weight"].shape[0]
    for impl in (1, 0):
        m = ChromeGCN(128, 128, nclass, 0.0, True, 2)
        m.load_state_dict(sd)
        m = m.to(_dev()).train()
        m.gemm_impl = impl
        eng = ChromosomeEngine(m, 2)
        loss = torch.zeros(1, device=_dev())
        xg = torch.empty(n, 2, 128, device=_dev())
        out, _ = eng.run(g, eng.pack(torch.from_numpy(z["x_f"]).to(_dev()), torch.from_numpy(z["x_r"]).to(_dev())),
                         torch.from_numpy(z["target"]).to(_dev()), None, loss, train=True, input_grad=xg)
        assert ogcn.max_rel(out.mean(1).cpu(), torch.from_numpy(z["f32.pred"])) <= 1e-5
        assert abs(loss.item() - float(z["f32.loss"])) <= 1e-5 * abs(float(z["f32.loss"]))
        for k, p in m.named_parameters():
            ref64 = torch.from_numpy(z["f64.grad." + k])
            own = ogcn.max_rel(torch.from_numpy(z["f32.grad." + k]), ref64)
            assert ogcn.max_rel(p.grad.cpu(), ref64) <= max(1e-5, 3 * own), (impl, k)


def _elementwise_rel(got, ref, floor):
    """max over elements of |got - ref| / max(|ref|, floor * max|ref|): the element-wise companion of `max_rel` (VERDICT
    r01: max-norm alone hides errors on small entries).  The floor is needed because every output here is a 128-term
    fp32 dot product: its absolute rounding error (~1e-6 of the row's scale, the fp32 reference's as much as ours) is
    unrelated to how close to zero the sum happens to land."""
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    denom = torch.clamp(ref.abs(), min=floor * float(ref.abs().max()))
    return float(((got - ref).abs() / denom).max())


@pytest.mark.parametrize("strands", [1, 2])
@pytest.mark.parametrize("mode", ["plain", "gate_off", "dropout", "stats", "hubs"])
def test_fused_layer_fwd_matches_fp64(strands, mode):
    """cgcn_gcn_layer_fwd (one kernel: gather -> tcgen05 3xTF32 -> gate epilogue) against the layer of
    models/ChromeModels.py:37-40 + models/SubLayers.py:42-52 in fp64 torch on the same inputs; the dropout case feeds
    the oracle the kernel's own keep-mask (cgcn_dropout_mask)."""
    from chromegcn_b200 import ops
    n = 2777 if mode != "hubs" else 1500                 # not a multiple of the 64-row tile; several CTAs
    ip, ix = _random_pattern(n, 14, 11, hub=900 if mode == "hubs" else None)     # row 7 of "hubs" spans 29 segments
    g = _graph(ip, ix)
    gen = torch.Generator().manual_seed(3)
    shape = (n, 128) if strands == 1 else (n, 2, 128)
    x = torch.randn(*shape, generator=gen)
    w = torch.randn(128, 128, generator=gen) * 0.12
    b = torch.randn(128, generator=gen) * 0.1
    wg = torch.randn(128, generator=gen) * 0.3
    bg = torch.randn(1, generator=gen) * 0.1
    p = 0.3 if mode == "dropout" else 0.0
    seed, step, site = 1234567, 9, 0
    xo, z, gate, sx, stats = ops.gcn_layer_fwd(g, x.to(_dev()), w.to(_dev()), b.to(_dev()), wg.to(_dev()), bg.to(_dev()),
                                               gate_off=(mode == "gate_off"), dropout_p=p, seed=seed, step=step, site=site,
                                               with_stats=(mode == "stats"))
    a = ogcn.coo_adjacency(ip, ix, torch.float64).coalesce()
    xd = x.double().reshape(n, -1)
    pat = torch.sparse_coo_tensor(a.indices(), torch.ones_like(a.values()), a.shape)
    sx_ref = torch.sparse.mm(pat, xd).reshape(shape)
    ax = torch.sparse.mm(a, xd).reshape(shape)
    z_ref = torch.tanh(ax @ w.double() + b.double())
    g_ref = torch.ones(*shape[:-1], 1, dtype=torch.float64) if mode == "gate_off" else torch.sigmoid(z_ref @ wg.double()[:, None] + bg.double())
    xo_ref = (1 - g_ref) * x.double() + g_ref * z_ref
    if p > 0:
        mask = ops.dropout_mask(n, strands, 128, p, seed, step, site, _dev()).cpu().double().reshape(shape)
        assert 0.6 < float((mask > 0).double().mean()) < 0.8
        xo_ref = xo_ref * mask
    assert ogcn.max_rel(sx.cpu(), sx_ref) <= 2e-6
    assert ogcn.max_rel(z.cpu(), z_ref) <= 1e-5 and _elementwise_rel(z, z_ref, 0.05) <= 1e-4
    assert ogcn.max_rel(gate.cpu().reshape(g_ref.shape), g_ref) <= 1e-5
    assert ogcn.max_rel(xo.cpu(), xo_ref) <= 1e-5 and _elementwise_rel(xo, xo_ref, 0.05) <= 1e-4
    if mode == "stats":
        r = torch.relu(xo_ref).reshape(n, strands, 128)
        got = stats.double().sum(0).cpu().reshape(2, strands, 128)
        assert ogcn.max_rel(got[0], r.sum(0)) <= 1e-5 and ogcn.max_rel(got[1], (r * r).sum(0)) <= 1e-5
