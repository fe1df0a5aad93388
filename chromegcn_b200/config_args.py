"""The reference's command line for the GCN stage (config_args.py:4-54, 57-143): the same single-dash
flags with the same defaults, and the same derived fields (`model_name`, `graph_root`, `dataset`,
forced `batch_size = 512` outside pre-training).  One deliberate difference: the
reference hard-codes `graph_root` to a cluster path (config_args.py:62); here it derives from
`-dataroot` like `opt.dataset` does."""
from __future__ import annotations

import os.path as path


def get_args(parser, argv=None):
    a = parser.add_argument
    a('-dataroot', type=str, default='./processed_data/')
    a('-results_dir', type=str, default='./results/')
    a('-cell_type', type=str, default='GM12878')
    a('-window_size', type=str, default='1000')
    a('-epochs', type=int, default=100)
    a('-batch_size', type=int, default=64)
    a('-test_batch_size', type=int, default=-1)
    a('-d_model', type=int, default=128)
    a('-optim', type=str, choices=['adam', 'sgd'], default='adam')
    a('-optim2', type=str, choices=['adam', 'sgd'], default='adam')
    a('-lr', type=float, default=0.0002)
    a('-lr2', type=float, default=0.002)
    a('-weight_decay', type=float, default=5e-5)
    a('-lr_decay', type=float, default=0)
    a('-lr_step_size', type=int, default=1)
    a('-lr_decay2', type=float, default=0)
    a('-lr_step_size2', type=int, default=100)
    a('-dropout', type=float, default=0.1)
    a('-gcn_dropout', type=float, default=0.2)
    a('-save_mode', type=str, choices=['all', 'best'], default='best')
    a('-window_model', type=str, choices=['deepsea', 'expecto', 'danq'], default='expecto')
    a('-loss', type=str, choices=['ce'], default='ce')
    a('-br_threshold', type=float, default=0.5)
    a('-no_cuda', action='store_true')
    a('-shuffle_train', action='store_true')
    a('-pretrain', action='store_true')
    a('-viz', action='store_true')
    a('-gpu_id', type=int, default=-1)
    a('-small', action='store_true')
    a('-summarize_data', action='store_true')
    a('-overwrite', action='store_true')
    a('-test_only', action='store_true')
    a('-load_pretrained', action='store_true')
    a('-seq_length', type=int, default=2000)
    a('-gcn_layers', type=int, default=2)
    a('-save_feats', action='store_true')
    a('-saved_model', type=str, default='')
    a('-A_saliency', action='store_true')
    a('-chrome_model', type=str, choices=['gcn', 'rnn'], default='gcn')
    a('-adj_type', type=str, choices=['constant', 'hic', 'both', 'random', 'none', ''], default='hic')
    a('-hicnorm', type=str, choices=['KR', 'VC', 'SQRTVC', ''], default='SQRTVC')
    a('-hicsize', type=str, choices=['125000', '250000', '500000', '1000000'], default='1000000')
    a('-gate', action='store_true')
    a('-load_gcn', action='store_true')
    a('-noeye', action='store_true')
    a('-name', type=str, default=None)
    a('-name2', type=str, default=None)
    # extension (not in the reference): keep features on the GPU across epochs
    a('-cache_features_on_device', action='store_true')
    return parser.parse_args(argv)


def config_args(opt):
    if opt.test_batch_size <= 0:
        opt.test_batch_size = opt.batch_size
    opt.graph_root = path.join(opt.dataroot, opt.cell_type, opt.window_size, 'hic')
    opt.dec_dropout = opt.dropout
    opt.drop_last = not opt.test_only
    name = 'graph.' + opt.window_model + '.' + str(opt.d_model) + '.bsz_' + str(opt.batch_size) + '.loss_' + str(opt.loss)
    name += '.' + str(opt.optim) + '.lr_' + str(opt.lr).split('.')[1]
    if opt.lr_decay > 0:
        name += '.decay_' + str(opt.lr_decay).replace('.', '') + '_' + str(opt.lr_step_size)
    name += '.drop_' + ("%.2f" % opt.dropout).split('.')[1] + '_' + ("%.2f" % opt.dec_dropout).split('.')[1]
    if opt.name:
        name += '.' + str(opt.name)
    if opt.save_feats:
        opt.pretrain, opt.shuffle_train, opt.epochs = False, False, 1
    elif opt.load_pretrained:
        name += '.finetune' + '.lr2_' + str(opt.lr2).split('.')[1] + '.gcndrop_' + ("%.2f" % opt.gcn_dropout).split('.')[1]
        name += '.' + str(opt.optim2) + '.' + str(opt.chrome_model) + '.layers_' + str(opt.gcn_layers)
        if opt.chrome_model == 'gcn' and opt.gate:
            name += '.gate'
        if opt.chrome_model == 'gcn':
            name += '.adj_' + opt.adj_type
            if opt.adj_type in ('hic', 'both'):
                name += '.norm_' + opt.hicnorm
            if opt.noeye:
                name += '.noeye'
        if opt.lr_decay2 > 0:
            name += '.decay_' + str(opt.lr_decay2).replace('.', '') + '_' + str(opt.lr_step_size2)
        if opt.name2 is not None:
            name += '.' + opt.name2
    opt.model_name = path.join(opt.results_dir, opt.cell_type, name)
    opt.dataset = path.join(opt.dataroot, opt.cell_type, opt.window_size)
    opt.cuda = not opt.no_cuda
    opt.d_word_vec = opt.d_model
    opt.data = path.join(opt.dataset, 'train_valid_test_small.pt' if opt.small else 'train_valid_test.pt')
    if opt.load_gcn:
        opt.model_name += '.load_gcn'
    if not opt.pretrain:
        opt.batch_size = 512
        opt.test_batch_size = 512
    opt.src_vocab_size = 5
    return opt
