import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from chromegcn_b200.chrome_models import ChromeGCN
from chromegcn_b200.engine import ChromosomeEngine
from chromegcn_b200.graph import HiCGraph
from chromegcn_b200.optim import FlatSGD
from oracle import gcn as ogcn
dev = torch.device("cuda", 0)
z = np.load("tests/golden/finetune.npz")
nclass = int(z["nclass"])
sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd0.")}
def data(c):
    ip, ix = z[c + ".indptr"], z[c + ".indices"]
    return ip, ix, [torch.from_numpy(z["%s.%s" % (c, k)]) for k in ("forward", "backward", "target")]
def oracle_grads(state, c):
    o = ogcn.ChromeGCNOracle(128, 128, nclass, 0.0, True, 2); o.load_state_dict(state); o = o.double().train()
    ip, ix, (xf, xr, t) = data(c)
    ogcn.chromosome_step(o, xf.double(), xr.double(), t.double(), ogcn.coo_adjacency(ip, ix, torch.float64), None, True)
    return {k: p.grad for k, p in o.named_parameters()}
for impl in (0, 1):
    for fresh_engine in (False, True):
        m = ChromeGCN(128, 128, nclass, 0.0, True, 2); m.load_state_dict(sd); m = m.to(dev).train(); m.gemm_impl = impl
        eng = ChromosomeEngine(m, 2); opt = FlatSGD(m, lr=0.25)
        for step, c in enumerate(("chr1", "chr2")):
            ip, ix, (xf, xr, t) = data(c)
            g = HiCGraph.from_csr_pattern(ip, ix, dev)
            state = {k: v.detach().cpu().clone() for k, v in m.state_dict().items()}
            want = oracle_grads(state, c)
            if fresh_engine: eng = ChromosomeEngine(m, 2)
            loss = torch.zeros(1, device=dev)
            eng.run(g, eng.pack(xf.to(dev), xr.to(dev)), t.to(dev), None, loss, train=True)
            errs = {k: ogcn.max_rel(p.grad.cpu(), want[k]) for k, p in m.named_parameters()}
            worst = max(errs, key=errs.get)
            print("impl %d fresh_engine %d step %d %s: worst grad err %s %.2e ; GC1.weight %.2e" % (impl, fresh_engine, step, c, worst, errs[worst], errs["GC1.weight"]))
            opt.step()
