"""GPU: the reference-facing drop-ins end to end -- create_graph() on Juicer-format text files, the
epoch driver, the CLI namespace -- on small synthetic inputs, checked against the golden vectors / oracle."""
import argparse
import os
import pickle

import numpy as np
import pytest
import torch

from oracle import adjacency as oadj

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
CHROMS = ["chr1", "chr2", "chr3", "chr22"]


@pytest.mark.parametrize("fname", ["adjacency_SQRTVC_1200.npz", "adjacency_none_1200.npz"])
def test_create_graph_from_text_files_matches_reference(tmp_path, fname):
    """Same files in, same pickles out as data/7create_graph_new.py (golden CSR minted from the reference)."""
    from chromegcn_b200 import synthetic
    from chromegcn_b200.create_graph import create_graph
    z = np.load(os.path.join(GOLDEN, fname))
    norm = str(z["norm_name"])
    hics = [synthetic.SyntheticHiC(c, z[c + "_windows"], z[c + "_bin1"], z[c + "_bin2"], z[c + "_val"], z[c + "_norm"], 1)
            for c in CHROMS]
    paths = synthetic.write_juicer_files(str(tmp_path), "GM12878", hics, norm_name=norm if norm else "SQRTVC")
    if norm == "":      # the golden arrays are already in `.sorted` file order: write them as that file
        for h in hics:
            d = os.path.join(paths["hic_root"], "GM12878_combined", "1kb_resolution_intrachromosomal", h.chrom, "MAPQGE30")
            os.replace(os.path.join(d, "%s_1kb.RAWobserved" % h.chrom), os.path.join(d, "%s_1kb.RAWobserved.sorted" % h.chrom))
    args = argparse.Namespace(output_root=paths["output_root"], use_all_windows=False, hic_root=paths["hic_root"],
                              cell_type="GM12878", resolution="1", hic_edges=int(z["hic_edges"]), norm=norm, chroms=CHROMS,
                              valid_chroms=["chr3"], test_chroms=["chr1"])
    create_graph(args)
    got = {}
    for split in ("train", "valid", "test"):
        f = os.path.join(paths["output_root"], "hic", "%s_graphs_%d_%snorm.pkl" % (split, int(z["hic_edges"]), norm))
        with open(f, "rb") as fp:
            got[split] = pickle.load(fp)
    assert sorted(got["valid"]) == ["chr3"] and sorted(got["test"]) == ["chr1"] and sorted(got["train"]) == ["chr2", "chr22"]
    for split in got:
        for c, csr in got[split].items():
            assert csr.dtype == np.float64 and np.all(csr.data == 1.0)
            assert np.array_equal(csr.indptr, z[c + "_indptr"]) and np.array_equal(csr.indices, z[c + "_indices"]), c


def test_run_model_two_epochs(tmp_path):
    """runner.run_model (runner.py:25-63): train / valid / test passes, metrics, CSV logs, best checkpoint in the
    reference's format; the loss must go down on a learnable synthetic task."""
    from scipy import sparse
    from chromegcn_b200 import synthetic
    from chromegcn_b200.chrome_models import ChromeGCN
    from chromegcn_b200.optim import get_optimizer
    from chromegcn_b200.runner import run_model
    from chromegcn_b200 import finetune as ft
    ft.clear_caches()
    dev = torch.device("cuda", 0)
    nclass, graphs, feats = 12, {}, {}
    gen = torch.Generator().manual_seed(0)
    wtrue = torch.randn(128, nclass, generator=gen)
    for c in ("chr20", "chr21", "chr22", "chr19"):
        h = synthetic.make_hic(c, hic_edges=6000, n_windows=600, n_bins=1500)
        ip, ix = oadj.build_adjacency_numpy(h.window_starts, h.bin1, h.bin2, h.val, h.norm, 1, 6000)
        n = ip.shape[0] - 1
        graphs[c] = sparse.csr_matrix((np.ones(ix.shape[0]), ix, ip), shape=(n, n))
        xf = torch.randn(n, 128, generator=gen)
        xr = xf + 0.1 * torch.randn(n, 128, generator=gen)
        feats[c] = {"forward": xf, "backward": xr, "target": ((xf @ wtrue) > 1.0).float()}
    root = tmp_path / "hic"
    root.mkdir()
    split_of = {"train": ["chr20", "chr21"], "valid": ["chr22"], "test": ["chr19"]}
    for s, cs in split_of.items():
        with open(root / ("%s_graphs_6000_SQRTVCnorm.pkl" % s), "wb") as fp:
            pickle.dump({c: graphs[c] for c in cs}, fp)
    opt = argparse.Namespace(adj_type="hic", graph_root=str(root), hicsize="6000", hicnorm="SQRTVC", optim="sgd", lr=0.25,
                             epochs=3, cell_type="GM12878", model_name=str(tmp_path / "model"), lr_decay2=0)
    torch.manual_seed(1)
    m = ChromeGCN(128, 128, nclass, 0.2, True, 2).to(dev)
    hist = run_model(None, m, {c: feats[c] for c in split_of["train"]}, {c: feats[c] for c in split_of["valid"]},
                     {c: feats[c] for c in split_of["test"]}, None, get_optimizer(m, opt), None, opt, None, None)
    assert len(hist) == 3 and hist[-1]["train_loss"] < hist[0]["train_loss"]
    assert 0.5 < hist[-1]["test_meanAUC"] <= 1.0
    ck = torch.load(tmp_path / "model" / "model.chkpt", weights_only=False)
    assert set(ck) == {"model", "settings", "epoch"} and "GC1.weight" in ck["model"]
    lines = open(tmp_path / "model" / "train.log").read().strip().splitlines()
    assert len(lines) == 3 and lines[0].startswith("1,") and len(lines[0].split(",")) == 6     # utils/evals.py:297-300: no header
    assert os.path.exists(tmp_path / "model" / "epochs" / "best_valid_preds_loss.pt")           # utils/evals.py:279-282
    assert os.path.exists(tmp_path / "model" / "epochs" / "best_test_targets_metrics.pt")      # :286-289
    # the checkpoint loads into the CPU oracle (== the reference's class layout)
    from oracle.gcn import ChromeGCNOracle
    ChromeGCNOracle(128, 128, nclass, 0.2, True, 2).load_state_dict(ck["model"])
    # the run above took its metrics from the GPU (cgcn_label_metrics on the device-resident probabilities);
    # the same run with the reference's sklearn route on the CPU copies must log the same numbers
    assert ft.DEVICE_OUTPUTS["test"][1] is not None
    ft.clear_caches()
    opt.device_metrics = False
    opt.model_name = str(tmp_path / "model_host_metrics")
    torch.manual_seed(1)
    m2 = ChromeGCN(128, 128, nclass, 0.2, True, 2).to(dev)
    hist2 = run_model(None, m2, {c: feats[c] for c in split_of["train"]}, {c: feats[c] for c in split_of["valid"]},
                      {c: feats[c] for c in split_of["test"]}, None, get_optimizer(m2, opt), None, opt, None, None)
    for a, b in zip(hist, hist2):
        assert a["train_loss"] == b["train_loss"] and a["test_loss"] == b["test_loss"]
        assert abs(a["test_meanAUC"] - b["test_meanAUC"]) <= 1e-9 and abs(a["test_meanAUPR"] - b["test_meanAUPR"]) <= 1e-9
    dev_log = open(tmp_path / "model" / "test.log").read().strip().splitlines()
    host_log = open(tmp_path / "model_host_metrics" / "test.log").read().strip().splitlines()
    assert len(dev_log) == 3
    for la, lb in zip(dev_log, host_log):
        va, vb = [float(x) for x in la.split(",")], [float(x) for x in lb.split(",")]
        assert max(abs(x - y) for x, y in zip(va, vb)) <= 1e-9
