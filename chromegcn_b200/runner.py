"""Drop-in epoch driver (reference: runner.py:10-63) for the GCN stage.

`run_epoch` / `run_model` keep the reference's signatures.  Only the fine-tuning branch exists here
(the CNN pre-training stage is out of scope): `opt.pretrain` / `opt.save_feats` raise.  Logging is the
reference's CSV line format (`utils/evals.py:297-300`) and checkpoint format (`:250-263`:
`{'model': state_dict, 'settings': opt, 'epoch': n}` in `<model_name>/model.chkpt`, written when the
validation score `meanAUPR + meanAUPR + meanFDR` (runner.py:46) is the best so far)."""
from __future__ import annotations

import os
import time

import torch

from . import finetune as _ft
from .finetune import finetune
from .metrics import compute_metrics as _compute_metrics_host
from .metrics import compute_metrics_device


def compute_metrics(preds, targs, loss, opt=None, elapsed=0.0, data_dict=None, cell_type=None, split=None):
    """utils/evals.py:86-120.  The split's probabilities and label bits are still on the GPU after `finetune()`
    (`finetune.DEVICE_OUTPUTS`), so the per-label AUROC / AUPR / FDR come from `cgcn_label_metrics` there; the
    sklearn route on the CPU copies remains for soft labels and for `opt.device_metrics = False`."""
    dev = _ft.DEVICE_OUTPUTS.get(split) if getattr(opt, "device_metrics", True) else None
    if dev is not None and dev[1] is not None and dev[0].shape[0] == preds.shape[0] and dev[0].shape[0] > 0:
        return compute_metrics_device(dev[0], dev[1], loss, opt, elapsed, data_dict, cell_type)
    return _compute_metrics_host(preds, targs, loss, opt, elapsed, data_dict, cell_type)


def run_epoch(WindowModel, ChromeModel, split_data, crit, optimizer, epoch, data_dict, opt, split):
    start = time.time()
    if getattr(opt, "pretrain", False) or getattr(opt, "save_feats", False):
        raise NotImplementedError("the CNN pre-training stage (pretrain.py) is outside this package")
    print("Finetune")
    pred, targ, loss = finetune(WindowModel, ChromeModel, split_data, crit, optimizer, epoch, data_dict, opt, split)
    elapsed = (time.time() - start) / 60
    print("\n({split}) elapse: {elapse:3.3f} min".format(split=split, elapse=elapsed))
    print("Loss: {loss:3.3f}".format(loss=loss))
    return pred, targ, loss, elapsed


class SaveLogger:
    """utils/evals.py:265-300: empty log files (no header line), `best_loss_epoch` = epoch of the lowest validation
    loss with its `epochs/best_*_loss.pt` predictions, `epochs/best_*_metrics.pt` for the best validation metric sum,
    and the `save_mode == 'best'` checkpoint rule of `save_model` (:250-263)."""

    def __init__(self, model_name):
        self.model_name = model_name
        self.best_valid_loss = float("inf")
        self.best_valid_metric = 0
        self.best_loss_epoch = 0
        os.makedirs(os.path.join(model_name, "epochs"), exist_ok=True)
        for f in ("train.log", "valid.log", "test.log"):
            open(os.path.join(model_name, f), "w").close()

    def log(self, file_name, epoch, loss, metrics):
        if metrics is None:
            return
        with open(os.path.join(self.model_name, file_name), "a") as fp:
            fp.write("%s,%s,%s,%s,%s,%s\n" % (epoch, loss, metrics.get("mAP", 0), metrics["meanAUC"], metrics["meanAUPR"],
                                              metrics["meanFDR"]))

    def _dump(self, tag, valid_preds, valid_targs, test_preds, test_targs):
        d = os.path.join(self.model_name, "epochs")
        torch.save(valid_preds, os.path.join(d, "best_valid_preds_%s.pt" % tag))
        torch.save(valid_targs, os.path.join(d, "best_valid_targets_%s.pt" % tag))
        torch.save(test_preds, os.path.join(d, "best_test_preds_%s.pt" % tag))
        torch.save(test_targs, os.path.join(d, "best_test_targets_%s.pt" % tag))

    def save(self, epoch, opt, ChromeModel, valid_loss, valid_metrics_sum, valid_metrics_sums, valid_preds, valid_targs,
             test_preds, test_targs):
        if valid_loss < self.best_valid_loss:                                  # utils/evals.py:276-283
            self.best_valid_loss = valid_loss
            self.best_loss_epoch = epoch
            self._dump("loss", valid_preds, valid_targs, test_preds, test_targs)
        if valid_metrics_sum > self.best_valid_metric:                         # :284-289
            self.best_valid_metric = valid_metrics_sum
            self._dump("metrics", valid_preds, valid_targs, test_preds, test_targs)
        # :291 (`not 'test' in self.model_name`: the reference tests the whole path string; only the run's own
        # directory name is examined here, so that a parent directory called e.g. "tests" does not disable saving)
        if "test" in os.path.basename(os.path.normpath(self.model_name)) or getattr(opt, "test_only", False):
            return
        ckpt = {"model": ChromeModel.state_dict(), "settings": opt, "epoch": epoch}
        if getattr(opt, "save_mode", "best") == "all":                         # :253-255
            torch.save(ckpt, os.path.join(self.model_name, "accu_{accu:3.3f}.chkpt".format(accu=100 * valid_metrics_sum)))
        elif valid_metrics_sums and valid_metrics_sum >= max(valid_metrics_sums):
            torch.save(ckpt, os.path.join(self.model_name, "model.chkpt"))
            print("[Info] The checkpoint file has been updated.")


def run_model(WindowModel, ChromeModel, train_data, valid_data, test_data, crit, optimizer, scheduler, opt, data_dict, logger):
    valid_metrics_sums = []
    save_logger = SaveLogger(opt.model_name)
    history = []
    for epoch in range(1, opt.epochs + 1):
        print("================= Epoch", epoch, "=================")
        if scheduler and getattr(opt, "lr_decay2", 0) > 0:
            scheduler.step()
        train_metrics, valid_metrics = None, None
        train_loss, valid_loss, valid_metrics_sum = 0, 0, 0
        valid_preds = valid_targs = None
        if not getattr(opt, "load_gcn", False) and not getattr(opt, "test_only", False):
            train_preds, train_targs, train_loss, elpsd = run_epoch(WindowModel, ChromeModel, train_data, crit, optimizer,
                                                                    epoch, data_dict, opt, "train")
            train_metrics = compute_metrics(train_preds, train_targs, train_loss, opt, elpsd, data_dict, opt.cell_type, "train")
            valid_preds, valid_targs, valid_loss, elpsd = run_epoch(WindowModel, ChromeModel, valid_data, crit, optimizer,
                                                                    epoch, data_dict, opt, "valid")
            valid_metrics = compute_metrics(valid_preds, valid_targs, valid_loss, opt, elpsd, data_dict, opt.cell_type, "valid")
            valid_metrics_sum = valid_metrics["meanAUPR"] + valid_metrics["meanAUPR"] + valid_metrics["meanFDR"]
            valid_metrics_sums += [valid_metrics_sum]
        test_preds, test_targs, test_loss, elpsd = run_epoch(WindowModel, ChromeModel, test_data, crit, optimizer, epoch,
                                                             data_dict, opt, "test")
        test_metrics = compute_metrics(test_preds, test_targs, test_loss, opt, elpsd, data_dict, opt.cell_type, "test")
        if logger is not None:
            logger.evaluate(train_metrics, valid_metrics, test_metrics, epoch, getattr(opt, "total_num_parameters", 0))
        save_logger.save(epoch, opt, ChromeModel, valid_loss, valid_metrics_sum, valid_metrics_sums, valid_preds, valid_targs,
                         test_preds, test_targs)
        save_logger.log("test.log", epoch, test_loss, test_metrics)
        save_logger.log("valid.log", epoch, valid_loss, valid_metrics)
        save_logger.log("train.log", epoch, train_loss, train_metrics)
        print("best loss epoch: " + str(save_logger.best_loss_epoch))
        print(opt.model_name)
        history.append({"epoch": epoch, "train_loss": train_loss, "valid_loss": valid_loss, "test_loss": test_loss,
                        "test_meanAUC": test_metrics["meanAUC"], "test_meanAUPR": test_metrics["meanAUPR"]})
    return history
