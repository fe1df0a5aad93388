import sys, os, time, argparse, pickle, tempfile, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scipy import sparse
from chromegcn_b200 import ops, synthetic, finetune as ft
from chromegcn_b200.chrome_models import ChromeGCN
from chromegcn_b200.optim import FlatSGD
dev = torch.device("cuda", 0)
arg = sys.argv[1] if len(sys.argv) > 1 else ""
chroms = (arg.split(",") if "chr" in arg else synthetic.WHOLE_GENOME[:int(arg)] if arg else synthetic.WHOLE_GENOME)
tmp = tempfile.mkdtemp(); gd = {}; feats = {}
for c in chroms:
    h = synthetic.make_hic(c)
    ip, ix = ops.adjacency_build(h.window_starts, h.bin1, h.bin2, h.val, h.norm, 1, 500000, dev)
    n = ip.shape[0] - 1
    gd[c] = sparse.csr_matrix((np.ones(ix.shape[0]), ix, ip), shape=(n, n))
    feats[c] = {k: v.pin_memory() for k, v in synthetic.make_features(c, n).items()}
pickle.dump(gd, open(os.path.join(tmp, "train_graphs_500000_SQRTVCnorm.pkl"), "wb"))
opt = argparse.Namespace(adj_type="hic", graph_root=tmp, hicsize="500000", hicnorm="SQRTVC")
m = ChromeGCN(128, 128, 103, 0.2, True, 2).to(dev); o = FlatSGD(m, lr=0.25)
for _ in range(2): ft.finetune(None, m, feats, None, o, 0, None, opt, "train")
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3): ft.finetune(None, m, feats, None, o, 0, None, opt, "train")
torch.cuda.synchronize()
print("finetune e2e ms/epoch: %.1f" % ((time.perf_counter() - t0) / 3 * 1e3))
pr = cProfile.Profile(); pr.enable()
ft.finetune(None, m, feats, None, o, 0, None, opt, "train")
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
