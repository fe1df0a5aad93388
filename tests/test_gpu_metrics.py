"""GPU: `cgcn_label_metrics` (per-label AUROC / AUPR / recall at 50 % FDR / average precision) against sklearn, the
library the reference calls (utils/metrics.py:25-26,148-183,238-253), on inputs with heavy score ties, single-class
labels, label counts that are not a multiple of 32 and row counts that span several scan tiles."""
import warnings

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _sklearn_reference(t, p, cutoff=0.5):
    from sklearn import metrics as skm
    c = t.shape[1]
    auc = np.full(c, np.nan)
    aupr, fdr, ap = np.zeros(c), np.zeros(c), np.zeros(c)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        for i in range(c):
            try:
                auc[i] = skm.roc_auc_score(t[:, i], p[:, i])
            except ValueError:
                pass
            precision, recall, _ = skm.precision_recall_curve(t[:, i], p[:, i], pos_label=1)
            aupr[i] = skm.auc(recall, precision)
            f = 1 - precision
            fdr[i] = recall[next(j for j, x in enumerate(f) if x <= cutoff)]
            ap[i] = skm.average_precision_score(t[:, i], p[:, i], pos_label=1)
    return auc, aupr, fdr, ap


@pytest.mark.parametrize("n,c,levels", [(10007, 37, 50), (4096, 103, 0), (4097, 5, 3), (1, 3, 0), (70001, 12, 1000)])
def test_label_metrics_match_sklearn(n, c, levels):
    from chromegcn_b200 import metrics, ops
    rng = np.random.default_rng(n + c)
    t = (rng.random((n, c)) < rng.uniform(0.02, 0.6, size=c)).astype(np.float32)
    p = rng.random((n, c)).astype(np.float32)
    p = np.clip(0.35 * t + 0.65 * p, 0, 1).astype(np.float32)          # informative scores
    if levels:
        p = (np.floor(p * levels) / levels).astype(np.float32)          # heavy ties
    if c >= 3:
        t[:, 0] = 0.0                                                    # no positives
        t[:, 1] = 1.0                                                    # no negatives
        p[:, 2] = 0.25                                                   # one threshold only
    dev = torch.device("cuda", 0)
    pd, td = torch.from_numpy(p).to(dev), torch.from_numpy(t).to(dev)
    got = metrics.label_metrics_device(pd, td)
    bits = ops.pack_targets(torch.from_numpy(t)).to(dev)
    got_b = metrics.label_metrics_device(pd, bits)
    for k in got:
        assert np.array_equal(got[k], got_b[k], equal_nan=True), k       # float labels and bit rows: same kernel path
    auc, aupr, fdr, ap = _sklearn_reference(t, p)
    assert np.array_equal(np.isnan(got["auroc"]), np.isnan(auc))
    ok = ~np.isnan(auc)
    assert np.abs(got["auroc"][ok] - auc[ok]).max(initial=0) <= 1e-9
    assert np.abs(got["aupr"] - aupr).max() <= 1e-9
    assert np.abs(got["fdr"] - fdr).max() <= 1e-12
    assert np.abs(got["ap"] - ap).max() <= 1e-9
    assert np.array_equal(got["npos"], t.sum(0).astype(np.float64))


def test_compute_metrics_device_matches_host_dictionary():
    """Same dictionary as the sklearn route of `compute_metrics` (utils/evals.py:86-120), strided device inputs."""
    from chromegcn_b200 import metrics
    rng = np.random.default_rng(5)
    n, c = 3001, 19
    t = (rng.random((n, c)) < 0.2).astype(np.float32)
    p = (0.5 * t + 0.5 * rng.random((n, c))).astype(np.float32)
    t[:, 4] = 0
    dev = torch.device("cuda", 0)
    wide = torch.zeros(n, c + 5, device=dev)
    wide[:, :c] = torch.from_numpy(p).to(dev)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        host = metrics.compute_metrics(torch.from_numpy(p.copy()), torch.from_numpy(t), 1.5)
    got = metrics.compute_metrics_device(wide[:, :c], torch.from_numpy(t).to(dev), 1.5)
    for k in ("meanAUC", "medianAUC", "varAUC", "meanAUPR", "medianAUPR", "meanFDR", "medianFDR", "mAP"):
        assert abs(got[k] - host[k]) <= 1e-9, k
    for k in ("allAUC", "allAUPR", "allFDR"):
        assert got[k].shape == host[k].shape and np.abs(got[k] - host[k]).max() <= 1e-9, k
