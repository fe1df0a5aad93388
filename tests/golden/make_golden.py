#!/usr/bin/env python
"""Mint the golden vectors under tests/golden/ from the LIVE reference.

Run in the build container only (needs /root/reference, CPU is enough):

    python tests/golden/make_golden.py

The reference (QData/ChromeGCN) ships no tests or known-answer vectors for the
chromosome-model path, so the parity pin is the reference itself executed here on
seeded inputs.  Inputs and outputs are committed as small .npz files so that the GPU
box (which has no /root/reference) can check both the oracle and the CUDA path.

What runs, unmodified, from the reference:
  data/7create_graph_new.py  create_graph(args)            -> adjacency_*.npz
  utils/util_methods.py      process_graph('hic', ...)     -> process_graph.npz
  models/ChromeModels.py     ChromeGCN forward + autograd  -> model_*.npz
  finetune.py                finetune(...) (torch.Tensor.cuda patched to identity because
                             finetune.py:30-36 hard-codes .cuda())  -> finetune.npz
"""
import argparse
import os
import pickle
import sys
import tempfile
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("CHROMEGCN_REFERENCE", "/root/reference")
sys.path.insert(0, REPO)
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "data"))
warnings.filterwarnings("ignore")

from chromegcn_b200 import synthetic  # noqa: E402


# --------------------------------------------------------------------------------------
# adversarial Hi-C input
# --------------------------------------------------------------------------------------
def adversarial_hic(chrom, n_bins, n_windows, n_rows, seed):
    rng = np.random.default_rng(seed)
    windows = np.sort(rng.choice(n_bins, size=n_windows, replace=False)).astype(np.int64) * 1000
    wset = windows // 1000
    rows = []
    for _ in range(n_rows):
        kind = rng.random()
        if kind < 0.70:        # window x window
            a, b = rng.choice(wset, 2, replace=True)
        elif kind < 0.85:      # at least one non-window bin
            a, b = rng.integers(0, n_bins, 2)
        else:                  # diagonal
            a = b = rng.choice(wset)
        rows.append((int(a), int(b)))
    rows = np.array(rows, dtype=np.int64)
    # duplicates of earlier keys (value overwritten later) and reversed keys
    dup = rows[rng.integers(0, n_rows // 2, n_rows // 6)]
    rev = rows[rng.integers(0, n_rows // 2, n_rows // 6)][:, ::-1]
    rows = np.concatenate([rows, dup, rev])
    rows = rows[rng.permutation(rows.shape[0])]
    vals = rng.choice(np.array([0.0, 1.0, 2.0, 3.0, 5.0, 8.0, 13.0]), size=rows.shape[0])
    norm = rng.choice(np.array([0.5, 1.0, 1.0, 2.0, 0.25]), size=n_bins).astype(np.float64)
    bad = rng.random(n_bins)
    norm[bad < 0.05] = np.nan
    norm[(bad >= 0.05) & (bad < 0.10)] = 0.0
    return synthetic.SyntheticHiC(chrom, windows, rows[:, 0] * 1000, rows[:, 1] * 1000, vals, norm, 1)


def run_reference_create_graph(hics, norm, hic_edges, valid=("chr3",), test=("chr1",)):
    ref7 = __import__("7create_graph_new")
    with tempfile.TemporaryDirectory() as tmp:
        paths = synthetic.write_juicer_files(tmp, "GM12878", hics, norm_name=norm if norm else "SQRTVC",
                                             write_sorted=(norm == ""))
        args = argparse.Namespace(output_root=paths["output_root"], use_all_windows=False,
                                  hic_root=paths["hic_root"], cell_type="GM12878", resolution="1",
                                  hic_edges=hic_edges, norm=norm, chroms=[h.chrom for h in hics],
                                  valid_chroms=list(valid), test_chroms=list(test), residuals=[0])
        ref7.tqdm = lambda it, **kw: it
        ref7.create_graph(args)
        out = {}
        for split in ("train", "valid", "test"):
            f = os.path.join(paths["output_root"], "hic", "%s_graphs_%d_%snorm.pkl" % (split, hic_edges, norm))
            with open(f, "rb") as fp:
                out.update(pickle.load(fp))
    return out


def golden_adjacency():
    from oracle import adjacency as oadj
    hics = [adversarial_hic("chr1", 420, 211, 2600, 11), adversarial_hic("chr2", 300, 157, 1800, 12),
            adversarial_hic("chr3", 64, 40, 300, 13)]
    # a realistic, distance-decay one too (ascending order, unique keys)
    hics.append(synthetic.make_hic("chr22", hic_edges=3000, n_windows=500, n_bins=1300))
    for norm, hic_edges in (("SQRTVC", 1200), ("", 1200), ("SQRTVC", 41), ("SQRTVC", 10 ** 6)):
        graphs = run_reference_create_graph(hics, norm, hic_edges)
        pack = {"hic_edges": np.int64(hic_edges), "norm_name": np.array(norm)}
        for h in hics:
            csr = graphs[h.chrom]
            assert csr.dtype == np.float64 and np.all(csr.data == 1.0)
            csr.sort_indices()
            c = h.chrom
            if norm == "":
                order = np.argsort(-h.val, kind="stable")      # the `.sorted` file the reference read
                b1, b2, v = h.bin1[order], h.bin2[order], h.val[order]
            else:
                b1, b2, v = h.bin1, h.bin2, h.val
            pack.update({c + "_windows": h.window_starts, c + "_bin1": b1, c + "_bin2": b2, c + "_val": v,
                         c + "_norm": h.norm, c + "_indptr": csr.indptr.astype(np.int32),
                         c + "_indices": csr.indices.astype(np.int32)})
            nrm = h.norm if norm else None
            for fn in (oadj.build_adjacency_loops, oadj.build_adjacency_numpy):
                ip, ix = fn(h.window_starts, b1, b2, v, nrm, 1, hic_edges)
                assert np.array_equal(ip, csr.indptr) and np.array_equal(ix, csr.indices), (fn.__name__, norm, c)
        name = "adjacency_%s_%d.npz" % (norm if norm else "none", hic_edges)
        np.savez_compressed(os.path.join(HERE, name), **pack)
        print("wrote", name, {h.chrom: int(graphs[h.chrom].nnz) for h in hics})
    return hics, run_reference_create_graph(hics, "SQRTVC", 1200)


def golden_process_graph(hics, graphs):
    from utils import util_methods as ref_um
    from oracle import adjacency as oadj
    pack = {}
    for h in hics:
        csr = graphs[h.chrom]
        t = ref_um.process_graph("hic", graphs, csr.shape[0], h.chrom)
        idx = t._indices().numpy()
        val = t._values().numpy()
        assert val.dtype == np.float32 and idx.dtype == np.int64
        r, c, v = oadj.normalize_hic(csr.indptr, csr.indices)
        assert np.array_equal(r, idx[0]) and np.array_equal(c, idx[1]) and np.array_equal(v, val)
        pack.update({h.chrom + "_indptr": csr.indptr.astype(np.int32), h.chrom + "_indices": csr.indices.astype(np.int32),
                     h.chrom + "_coo_rows": idx[0], h.chrom + "_coo_cols": idx[1], h.chrom + "_coo_vals": val})
    np.savez_compressed(os.path.join(HERE, "process_graph.npz"), **pack)
    print("wrote process_graph.npz")


def golden_adj_types(hics, graphs):
    """process_graph for adj_type constant / both / none, and a ChromeGCN step on the 'both' graph."""
    from models.ChromeModels import ChromeGCN
    from utils import util_methods as ref_um
    from oracle import adjacency as oadj
    from oracle import gcn as ogcn
    h = hics[1]
    csr = graphs[h.chrom]
    n = csr.shape[0]
    pack = {"indptr": csr.indptr.astype(np.int32), "indices": csr.indices.astype(np.int32), "n": np.int64(n)}
    for t in ("constant", "both", "none"):
        ten = ref_um.process_graph(t, graphs, n, h.chrom).coalesce()
        idx, val = ten.indices().numpy(), ten.values().numpy()
        r, c, v = oadj.process_graph_general(t, csr.indptr, csr.indices, n)
        assert np.array_equal(r, idx[0]) and np.array_equal(c, idx[1]) and np.array_equal(v, val), t
        pack.update({t + "_rows": idx[0], t + "_cols": idx[1], t + "_vals": val})
    d, nclass = 128, 29
    g = torch.Generator().manual_seed(41)
    x_f, x_r = torch.randn(n, d, generator=g), torch.randn(n, d, generator=g)
    tgt = (torch.rand(n, nclass, generator=g) < 0.1).float()
    torch.manual_seed(6)
    model = ogcn.stress_init_(ChromeGCN(d, d, nclass, 0.0, True, 2), seed=11)
    pack.update(state_to_np(model.state_dict(), "sd0."))
    pack.update({"x_f": x_f.numpy(), "x_r": x_r.numpy(), "target": tgt.numpy()})
    adj = ref_um.process_graph("both", graphs, n, h.chrom)
    for dt_name, dt in (("f32", torch.float32), ("f64", torch.float64)):
        m = ChromeGCN(d, d, nclass, 0.0, True, 2)
        m.load_state_dict(model.state_dict())
        m = m.to(dt).train()
        _, pf, _, _ = m(x_f.to(dt), adj.to(dt), None)
        _, pr, _, _ = m(x_r.to(dt), adj.to(dt), None)
        pred = (pf + pr) / 2
        loss = torch.nn.functional.binary_cross_entropy_with_logits(pred, tgt.to(dt))
        loss.backward()
        pack["%s.pred" % dt_name] = pred.detach().numpy()
        pack["%s.loss" % dt_name] = np.array(loss.item())
        for k, p in m.named_parameters():
            pack["%s.grad.%s" % (dt_name, k)] = p.grad.numpy()
    np.savez_compressed(os.path.join(HERE, "adj_types.npz"), **pack)
    print("wrote adj_types.npz")


def state_to_np(sd, prefix):
    return {prefix + k: v.detach().cpu().numpy() for k, v in sd.items()}


def golden_model(hics, graphs):
    from models.ChromeModels import ChromeGCN
    from utils import util_methods as ref_um
    from oracle import gcn as ogcn
    d, nclass = 128, 103
    for tag, layers, stress in (("l2_ref", 2, False), ("l2_stress", 2, True), ("l1_stress", 1, True)):
        h = hics[0]
        csr = graphs[h.chrom]
        n = csr.shape[0]
        adj = ref_um.process_graph("hic", graphs, n, h.chrom)
        g = torch.Generator().manual_seed(99 + layers)
        x_f = torch.randn(n, d, generator=g)
        x_r = torch.randn(n, d, generator=g)
        tgt = (torch.rand(n, nclass, generator=g) < 0.1).float()
        torch.manual_seed(5)
        model = ChromeGCN(d, d, nclass, 0.0, True, layers)
        if stress:
            ogcn.stress_init_(model)
        pack = {"indptr": csr.indptr.astype(np.int32), "indices": csr.indices.astype(np.int32),
                "x_f": x_f.numpy(), "x_r": x_r.numpy(), "target": tgt.numpy(), "layers": np.int64(layers)}
        pack.update(state_to_np(model.state_dict(), "sd0."))
        for dt_name, dt in (("f32", torch.float32), ("f64", torch.float64)):
            torch.manual_seed(5)
            m = ChromeGCN(d, d, nclass, 0.0, True, layers)
            m.load_state_dict(model.state_dict())
            m = m.to(dt)
            a = adj.to(dt)
            # eval-mode forward (running stats 0 / 1)
            m.eval()
            with torch.no_grad():
                _, out_e, (g1, g2), _ = m(x_f.to(dt), a, None)
            pack["%s.eval.out_f" % dt_name] = out_e.numpy()
            pack["%s.eval.g1_f" % dt_name] = g1.numpy()
            if g2 is not None:
                pack["%s.eval.g2_f" % dt_name] = g2.numpy()
            # one train-mode step exactly as finetune.py:33-49 (dropout p = 0)
            m.train()
            xf = x_f.to(dt).clone().requires_grad_(True)
            xr = x_r.to(dt).clone().requires_grad_(True)
            _, pf, (g1f, g2f), _ = m(xf, a, None)
            _, pr, (g1r, g2r), _ = m(xr, a, None)
            pred = (pf + pr) / 2
            loss = torch.nn.functional.binary_cross_entropy_with_logits(pred, tgt.to(dt))
            loss.backward()
            pack["%s.train.pred" % dt_name] = pred.detach().numpy()
            pack["%s.train.loss" % dt_name] = np.array(loss.item())
            pack["%s.train.g1_f" % dt_name] = g1f.detach().numpy()
            pack["%s.train.g1_r" % dt_name] = g1r.detach().numpy()
            if g2f is not None:
                pack["%s.train.g2_f" % dt_name] = g2f.detach().numpy()
                pack["%s.train.g2_r" % dt_name] = g2r.detach().numpy()
            pack["%s.train.xgrad_f" % dt_name] = xf.grad.numpy()
            pack["%s.train.xgrad_r" % dt_name] = xr.grad.numpy()
            for k, p in m.named_parameters():
                pack["%s.grad.%s" % (dt_name, k)] = p.grad.numpy()
            pack.update(state_to_np({k: v for k, v in m.state_dict().items() if "running" in k or "num_batches" in k},
                                    "%s.after." % dt_name))
            if dt_name == "f32":
                # the oracle restatement must agree with the live reference to the last bit here
                om = ogcn.ChromeGCNOracle(d, d, nclass, 0.0, True, layers)
                om.load_state_dict(model.state_dict())
                om.train()
                oadjt = ogcn.coo_adjacency(csr.indptr, csr.indices)
                lo, _, po, ex = ogcn.chromosome_step(om, x_f, x_r, tgt, oadjt, None, True, input_grads=True)
                assert torch.equal(po, pred.detach()), "oracle forward differs from the live reference"
                for (k, p), (k2, p2) in zip(m.named_parameters(), om.named_parameters()):
                    assert k == k2 and torch.equal(p.grad, p2.grad), k
        np.savez_compressed(os.path.join(HERE, "model_%s.npz" % tag), **pack)
        print("wrote model_%s.npz" % tag, "loss", float(pack["f32.train.loss"]))


def golden_finetune(hics, graphs):
    """Three epochs of the reference's own finetune() over two training chromosomes and one
    validation chromosome, SGD lr 0.25 (README.md:45 recipe) with gcn_dropout 0."""
    import finetune as ref_ft
    from models.ChromeModels import ChromeGCN
    from utils import util_methods as ref_um
    from oracle import gcn as ogcn
    ref_ft.tqdm = lambda it, **kw: it
    orig_cuda = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        d, nclass = 128, 37
        train_chroms, valid_chroms = ["chr1", "chr2"], ["chr3"]
        feats = {}
        for h in hics[:3]:
            g = torch.Generator().manual_seed(300 + synthetic.chrom_index(h.chrom))
            n = graphs[h.chrom].shape[0]
            feats[h.chrom] = {"forward": torch.randn(n, d, generator=g), "backward": torch.randn(n, d, generator=g),
                              "target": (torch.rand(n, nclass, generator=g) < 0.15).float()}
        torch.manual_seed(17)
        model = ChromeGCN(d, d, nclass, 0.0, True, 2)
        ogcn.stress_init_(model, seed=3)
        sd0 = {k: v.clone() for k, v in model.state_dict().items()}
        omodel = ogcn.ChromeGCNOracle(d, d, nclass, 0.0, True, 2)
        omodel.load_state_dict(sd0)
        with tempfile.TemporaryDirectory() as tmp:
            with open(os.path.join(tmp, "train_graphs_1200_SQRTVCnorm.pkl"), "wb") as fp:
                pickle.dump({c: graphs[c] for c in train_chroms}, fp)
            with open(os.path.join(tmp, "valid_graphs_1200_SQRTVCnorm.pkl"), "wb") as fp:
                pickle.dump({c: graphs[c] for c in valid_chroms}, fp)
            opt = argparse.Namespace(adj_type="hic", graph_root=tmp, hicsize="1200", hicnorm="SQRTVC",
                                     optim="sgd", lr=0.25)
            optimizer = ref_um.get_optimizer(model, opt)
            oopt = ogcn.make_optimizer(omodel, "sgd", 0.25)
            pack = {"nclass": np.int64(nclass)}
            for c in train_chroms + valid_chroms:
                pack[c + ".forward"] = feats[c]["forward"].numpy()
                pack[c + ".backward"] = feats[c]["backward"].numpy()
                pack[c + ".target"] = feats[c]["target"].numpy()
                pack[c + ".indptr"] = graphs[c].indptr.astype(np.int32)
                pack[c + ".indices"] = graphs[c].indices.astype(np.int32)
            pack.update(state_to_np(sd0, "sd0."))
            train_d = {c: feats[c] for c in train_chroms}
            valid_d = {c: feats[c] for c in valid_chroms}
            og = {c: (graphs[c].indptr, graphs[c].indices) for c in graphs}
            for epoch in range(1, 4):
                p, t, l = ref_ft.finetune(None, model, train_d, None, optimizer, epoch, None, opt, "train")
                pv, tv, lv = ref_ft.finetune(None, model, valid_d, None, optimizer, epoch, None, opt, "valid")
                pack["epoch%d.train_loss" % epoch] = np.array(l)
                pack["epoch%d.valid_loss" % epoch] = np.array(lv)
                pack["epoch%d.train_preds" % epoch] = p.numpy()
                pack["epoch%d.valid_preds" % epoch] = pv.numpy()
                po, to, lo = ogcn.finetune_epoch(omodel, train_d, og, oopt, "train")
                pvo, tvo, lvo = ogcn.finetune_epoch(omodel, valid_d, og, oopt, "valid")
                assert torch.equal(po, p) and torch.equal(pvo, pv) and lo == l and lvo == lv, "oracle finetune differs"
                print("epoch", epoch, "train loss", l, "valid loss", lv)
            pack.update(state_to_np(model.state_dict(), "sd3."))
        np.savez_compressed(os.path.join(HERE, "finetune.npz"), **pack)
        print("wrote finetune.npz")
    finally:
        torch.Tensor.cuda = orig_cuda


def main():
    torch.set_num_threads(1)         # deterministic reduction order for the fp32 vectors
    hics, graphs = golden_adjacency()
    golden_process_graph(hics, graphs)
    golden_adj_types(hics, graphs)
    golden_model(hics, graphs)
    golden_finetune(hics, graphs)


if __name__ == "__main__":
    main()
