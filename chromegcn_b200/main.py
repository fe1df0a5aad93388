"""GCN stage of the reference's `main.py` (main.py:15-112) on the CUDA path:

    python -m chromegcn_b200.main -load_pretrained -chrome_model gcn -gate -gcn_layers 2 \\
        -adj_type hic -hicnorm SQRTVC -hicsize 500000 -optim sgd -lr 0.25 -gcn_dropout 0.2 -epochs 1000

Reads the reference's files: `<model_name without .finetune*>/chrom_feature_dict_{train,valid,test}.pt`
(utils/util_methods.py:183-199), the graph pickles under `opt.graph_root`, and, when present, the CNN
checkpoint whose classifier / batch-norm weights seed the GCN head (main.py:74-81)."""
from __future__ import annotations

import argparse
import os

import torch

from .chrome_models import ChromeGCN
from .config_args import config_args, get_args
from .optim import get_optimizer
from .runner import run_model


def main(argv=None):
    opt = config_args(get_args(argparse.ArgumentParser(), argv))
    if opt.pretrain or opt.save_feats or opt.chrome_model != 'gcn':
        raise NotImplementedError("only the GCN fine-tuning stage (-load_pretrained -chrome_model gcn) is implemented")
    base = opt.model_name.split('.finetune')[0]
    load = lambda s: torch.load(os.path.join(base, 'chrom_feature_dict_%s.pt' % s), weights_only=False)
    train_data, valid_data, test_data = load('train'), load('valid'), load('test')
    nclass = next(iter(train_data.values()))['target'].shape[1]
    opt.tgt_vocab_size = nclass
    model = ChromeGCN(128, 128, nclass, opt.gcn_dropout, opt.gate, opt.gcn_layers)       # main.py:62
    ckpt = os.path.join(opt.model_name.replace('.load_gcn', ''), 'model.chkpt') if opt.load_gcn else os.path.join(base, 'model.chkpt')
    if not os.path.exists(ckpt) and not os.environ.get('CHROMEGCN_RANDOM_HEAD_INIT'):
        # main.py:66-73 loads unconditionally (torch.load raises); training on a random head is opt-in only
        raise FileNotFoundError("%s: the %s checkpoint is required (set CHROMEGCN_RANDOM_HEAD_INIT=1 to train from a "
                                "random head instead)" % (ckpt, 'GCN' if opt.load_gcn else 'CNN'))
    if os.path.exists(ckpt):
        sd = torch.load(ckpt, weights_only=False)['model']
        if opt.load_gcn:
            model.load_state_dict(sd)
        else:                                                                             # main.py:78-81
            pick = lambda suffix: next(v for k, v in sd.items() if k.endswith(suffix))
            model.out.weight.data = pick('classifier.weight')
            model.out.bias.data = pick('classifier.bias')
            model.batch_norm.weight.data = pick('model.batch_norm.weight')
            model.batch_norm.bias.data = pick('model.batch_norm.bias')
    model = model.cuda()
    optimizer = get_optimizer(model, opt)                                                # main.py:83
    scheduler = torch.optim.lr_scheduler.StepLR(optimizer, step_size=100, gamma=0.5)      # main.py:86
    return run_model(None, model, train_data, valid_data, test_data, None, optimizer, scheduler, opt, None, None)


if __name__ == "__main__":
    main()
