"""Evaluation used by `runner.run_model` (reference: utils/metrics.py:148-183,238-253 and
utils/evals.py:86-120).

Two routes to the same dictionary:
  * `compute_metrics` -- sklearn on CPU copies of the predictions, exactly like the reference; it is the
    AUROC / AUPR parity instrument of the tests;
  * `compute_metrics_device` -- `cgcn_label_metrics` (csrc/metrics.cu): one radix sort + one scan pass on the
    GPU over the `[sum N, nclass]` probability matrix that `finetune()` leaves on the device (SURVEY.md
    section 8(f) rank 4: per split per epoch sklearn takes tens of seconds on the whole genome, the device
    path milliseconds).  No CPU fallback inside it.
"""
from __future__ import annotations

import numpy as np


def label_metrics_device(preds, targets, fdr_cutoff: float = 0.5):
    """Per-label AUROC / AUPR / recall at the FDR cutoff / average precision / positives on the GPU.
    `preds` `[n, C]` CUDA fp32; `targets` `[n, C]` CUDA fp32 (non-zero = positive) or the int32 bit rows of
    `ops.pack_targets` `[n, ceil(C/32)]`.  Returns a dict of float64 numpy arrays of length C (`auroc` is NaN
    where the label has a single class)."""
    import torch
    from . import _lib
    lib = _lib.load()
    _lib.require_cuda(preds.device)
    if preds.dtype != torch.float32 or preds.dim() != 2 or preds.stride(1) != 1:
        preds = preds.float().contiguous()
    n, c = preds.shape
    bits = targets.dtype == torch.int32
    if bits:
        if tuple(targets.shape) != (n, (c + 31) // 32):
            raise ValueError("bit-packed targets must be [n, ceil(C/32)] int32")
        targets = targets.contiguous()
    else:
        if tuple(targets.shape) != (n, c):
            raise ValueError("targets must be [n, C]")
        if targets.dtype != torch.float32 or targets.stride(1) != 1:
            targets = targets.float().contiguous()
    if targets.device != preds.device:
        raise ValueError("preds and targets must be on the same device")
    with torch.cuda.device(preds.device):
        out = torch.empty(5, c, dtype=torch.float64, device=preds.device)
        need = lib.cgcn_label_metrics_workspace_bytes(n, c)
        ws = torch.empty(need, dtype=torch.uint8, device=preds.device)
        _lib.check(lib.cgcn_label_metrics(preds.data_ptr(), preds.stride(0), None if bits else targets.data_ptr(),
                                          0 if bits else targets.stride(0), targets.data_ptr() if bits else None, n, c,
                                          float(fdr_cutoff), out.data_ptr(), ws.data_ptr(), need, _lib.current_stream()),
                   "cgcn_label_metrics")
        host = out.cpu().numpy()
    return {"auroc": host[0], "aupr": host[1], "fdr": host[2], "ap": host[3], "npos": host[4]}


def compute_metrics_device(preds, targets, loss, opt=None, elapsed=0.0, data_dict=None, cell_type=None):
    """`compute_metrics` from device-resident predictions / labels (same keys, same label-skipping rules:
    AUROC only over labels with both classes, utils/metrics.py:243-247)."""
    m = label_metrics_device(preds, targets)
    auc_arr = m["auroc"][~np.isnan(m["auroc"])]
    aupr_arr, fdr_arr = m["aupr"], m["fdr"]
    mean = lambda a: float(np.mean(a)) if a.size else float("nan")
    med = lambda a: float(np.median(a)) if a.size else float("nan")
    var = lambda a: float(np.var(a)) if a.size else float("nan")
    return {"loss": loss, "time": elapsed, "mAP": mean(m["ap"]), "meanAUC": mean(auc_arr), "medianAUC": med(auc_arr),
            "varAUC": var(auc_arr), "allAUC": auc_arr, "meanAUPR": mean(aupr_arr), "medianAUPR": med(aupr_arr),
            "varAUPR": var(aupr_arr), "allAUPR": aupr_arr, "meanFDR": mean(fdr_arr), "medianFDR": med(fdr_arr),
            "varFDR": var(fdr_arr), "allFDR": fdr_arr}


def auroc(all_targets, all_predictions):
    """utils/metrics.py:238-253."""
    from sklearn import metrics as skm
    vals = []
    for i in range(all_targets.shape[1]):
        try:
            v = skm.roc_auc_score(all_targets[:, i], all_predictions[:, i])
            if not np.isnan(v):       # single-class label: sklearn of the reference's era raises (skipped at :245-246),
                vals.append(v)        # sklearn >= 1.7 warns and returns NaN -- skipped here as well
        except ValueError:
            pass
    vals = np.array(vals)
    return float(np.mean(vals)), float(np.median(vals)), float(np.var(vals)), vals


def aupr(all_targets, all_predictions):
    """utils/metrics.py:168-183."""
    from sklearn import metrics as skm
    vals = []
    for i in range(all_targets.shape[1]):
        try:
            precision, recall, _ = skm.precision_recall_curve(all_targets[:, i], all_predictions[:, i], pos_label=1)
            v = skm.auc(recall, precision)
            if not np.isnan(v):
                vals.append(np.nan_to_num(v))
        except Exception:
            pass
    vals = np.array(vals)
    return float(np.mean(vals)), float(np.median(vals)), float(np.var(vals)), vals


def fdr(all_targets, all_predictions, fdr_cutoff=0.5):
    """Recall at 50 % FDR, utils/metrics.py:148-165."""
    from sklearn import metrics as skm
    vals = []
    for i in range(all_targets.shape[1]):
        try:
            precision, recall, _ = skm.precision_recall_curve(all_targets[:, i], all_predictions[:, i], pos_label=1)
            f = 1 - precision
            idx = np.where(f <= fdr_cutoff)[0]
            vals.append(float(recall[idx[0]]) if idx.size else 0.0)
        except Exception:
            pass
    vals = np.array(vals)
    return float(np.mean(vals)), float(np.median(vals)), float(np.var(vals)), vals


def compute_metrics(all_predictions, all_targets, loss, opt=None, elapsed=0.0, data_dict=None, cell_type=None):
    """The dictionary `utils/evals.compute_metrics` returns (:86-120), without its side effects (it
    thresholds the predictions in place at `br_threshold`; here the inputs are left untouched)."""
    preds = all_predictions.numpy() if hasattr(all_predictions, "numpy") else np.asarray(all_predictions)
    targs = all_targets.numpy() if hasattr(all_targets, "numpy") else np.asarray(all_targets)
    mean_auc, median_auc, var_auc, auc_arr = auroc(targs, preds)
    mean_aupr, median_aupr, var_aupr, aupr_arr = aupr(targs, preds)
    mean_fdr, median_fdr, var_fdr, fdr_arr = fdr(targs, preds)
    try:                                      # utils/metrics.py:25-26
        from sklearn import metrics as skm
        m_ap = float(skm.average_precision_score(targs, preds, average="macro", pos_label=1))
    except Exception:
        m_ap = float("nan")
    return {"loss": loss, "time": elapsed, "mAP": m_ap, "meanAUC": mean_auc, "medianAUC": median_auc, "varAUC": var_auc,
            "allAUC": auc_arr, "meanAUPR": mean_aupr, "medianAUPR": median_aupr, "varAUPR": var_aupr,
            "allAUPR": aupr_arr, "meanFDR": mean_fdr, "medianFDR": median_fdr, "varFDR": var_fdr, "allFDR": fdr_arr}
