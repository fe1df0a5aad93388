"""Host-side evaluation used by `runner.run_model` (reference: utils/metrics.py:148-183,238-253 and
utils/evals.py:86-120).  sklearn on CPU copies of the predictions, like the reference; it is outside
the hot path and serves as the AUROC / AUPR parity instrument."""
from __future__ import annotations

import numpy as np


def auroc(all_targets, all_predictions):
    """utils/metrics.py:238-253."""
    from sklearn import metrics as skm
    vals = []
    for i in range(all_targets.shape[1]):
        try:
            vals.append(skm.roc_auc_score(all_targets[:, i], all_predictions[:, i]))
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
    return {"loss": loss, "time": elapsed, "meanAUC": mean_auc, "medianAUC": median_auc, "varAUC": var_auc,
            "allAUC": auc_arr, "meanAUPR": mean_aupr, "medianAUPR": median_aupr, "varAUPR": var_aupr,
            "allAUPR": aupr_arr, "meanFDR": mean_fdr, "medianFDR": median_fdr, "varFDR": var_fdr, "allFDR": fdr_arr}
