import sys, os, argparse, pickle, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from scipy import sparse
from chromegcn_b200.chrome_models import ChromeGCN
from chromegcn_b200 import finetune as ft
from chromegcn_b200.optim import get_optimizer
from oracle import gcn as ogcn
dev = torch.device("cuda", 0)
z = np.load("tests/golden/finetune.npz")
nclass = int(z["nclass"])
sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd0.")}
tmp = tempfile.mkdtemp()
graphs = {}
for c in ("chr1", "chr2", "chr3"):
    ip, ix = z[c + ".indptr"], z[c + ".indices"]; n = ip.shape[0] - 1
    graphs[c] = sparse.csr_matrix((np.ones(ix.shape[0]), ix, ip), shape=(n, n))
pickle.dump({c: graphs[c] for c in ("chr1", "chr2")}, open(os.path.join(tmp, "train_graphs_1200_SQRTVCnorm.pkl"), "wb"))
opt = argparse.Namespace(adj_type="hic", graph_root=tmp, hicsize="1200", hicnorm="SQRTVC", optim="sgd", lr=0.25)
feats = lambda cs: {c: {k: torch.from_numpy(z["%s.%s" % (c, k)]) for k in ("forward", "backward", "target")} for c in cs}
train_d = feats(["chr1", "chr2"])
og = {c: (z[c + ".indptr"], z[c + ".indices"]) for c in ("chr1", "chr2", "chr3")}
for nchrom in (2,):
    cs = ["chr1", "chr2"][:nchrom]
    o64 = ogcn.ChromeGCNOracle(128, 128, nclass, 0.0, True, 2); o64.load_state_dict(sd); o64 = o64.double()
    oopt = ogcn.make_optimizer(o64, "sgd", 0.25)
    ogcn.finetune_epoch(o64, {c: {k: v.double() for k, v in train_d[c].items()} for c in cs}, og, oopt, "train")
    o32 = ogcn.ChromeGCNOracle(128, 128, nclass, 0.0, True, 2); o32.load_state_dict(sd)
    o32opt = ogcn.make_optimizer(o32, "sgd", 0.25)
    ogcn.finetune_epoch(o32, {c: train_d[c] for c in cs}, og, o32opt, "train")
    for impl in (0,):
        m = ChromeGCN(128, 128, nclass, 0.0, True, 2); m.load_state_dict(sd); m = m.to(dev); m.gemm_impl = impl
        optimizer = get_optimizer(m, opt)
        ft.clear_caches()
        ft.finetune(None, m, {c: train_d[c] for c in cs}, None, optimizer, 1, None, opt, "train")
        print("after %d chromosome step(s), impl %d" % (nchrom, impl))
        for k, v in m.state_dict().items():
            if "num_batches" in k: continue
            ref = o64.state_dict()[k]
            if k in ("GC1.weight", "W2.bias", "GC2.weight"): print("   %-28s ours %.2e   torch-fp32 %.2e" % (k, ogcn.max_rel(v.float().cpu(), ref), ogcn.max_rel(o32.state_dict()[k], ref)))
