import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from chromegcn_b200.chrome_models import ChromeGCN
from chromegcn_b200.engine import ChromosomeEngine
from chromegcn_b200.graph import HiCGraph
from chromegcn_b200 import ops
from oracle import gcn as ogcn
dev = torch.device("cuda", 0)
z = np.load("tests/golden/finetune.npz")
nclass = int(z["nclass"])
sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd0.")}
o64 = ogcn.ChromeGCNOracle(128, 128, nclass, 0.0, True, 2); o64.load_state_dict(sd); o64 = o64.double().train()
for impl in (0, 1):
    m = ChromeGCN(128, 128, nclass, 0.0, True, 2); m.load_state_dict(sd); m = m.to(dev).train(); m.gemm_impl = impl
    eng = ChromosomeEngine(m, 2)
    for c in ("chr1", "chr2", "chr1", "chr3", "chr2"):
        ip, ix = z[c + ".indptr"], z[c + ".indices"]
        g = HiCGraph.from_csr_pattern(ip, ix, dev)
        xf, xr, t = (torch.from_numpy(z["%s.%s" % (c, k)]) for k in ("forward", "backward", "target"))
        o64.zero_grad()
        lo, _, pred_o, _ = ogcn.chromosome_step(o64, xf.double(), xr.double(), t.double(), ogcn.coo_adjacency(ip, ix, torch.float64), None, True)
        loss = torch.zeros(1, device=dev)
        out, _ = eng.run(g, eng.pack(xf.to(dev), xr.to(dev)), t.to(dev), None, loss, train=True)
        errs = {k: ogcn.max_rel(p.grad.cpu(), q.grad) for (k, p), (_, q) in zip(m.named_parameters(), o64.named_parameters())}
        worst = max(errs, key=errs.get)
        print("impl %d %s n=%d  logits err %.2e  worst grad %s %.2e" % (impl, c, xf.shape[0], ogcn.max_rel(out.mean(1).cpu(), pred_o), worst, errs[worst]))
