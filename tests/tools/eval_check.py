import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np, torch
from chromegcn_b200.chrome_models import ChromeGCN
from chromegcn_b200.engine import ChromosomeEngine
from chromegcn_b200.graph import HiCGraph
from oracle import gcn as ogcn
dev = torch.device("cuda", 0)
z = np.load("tests/golden/finetune.npz")
nclass = int(z["nclass"])
for which in ("sd0.", "sd3."):
    sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith(which)}
    o64 = ogcn.ChromeGCNOracle(128, 128, nclass, 0.0, True, 2); o64.load_state_dict(sd); o64 = o64.double().eval()
    for impl in (1, 0):
        m = ChromeGCN(128, 128, nclass, 0.0, True, 2); m.load_state_dict(sd); m = m.to(dev).eval(); m.gemm_impl = impl
        for c in ("chr3", "chr1"):
            ip, ix = z[c + ".indptr"], z[c + ".indices"]
            g = HiCGraph.from_csr_pattern(ip, ix, dev)
            xf, xr, t = (torch.from_numpy(z["%s.%s" % (c, k)]) for k in ("forward", "backward", "target"))
            with torch.no_grad():
                _, pf, _, _ = o64(xf.double(), ogcn.coo_adjacency(ip, ix, torch.float64))
                _, pr, _, _ = o64(xr.double(), ogcn.coo_adjacency(ip, ix, torch.float64))
            want = torch.sigmoid((pf + pr) / 2)
            eng = ChromosomeEngine(m, 2)
            probs = torch.empty(xf.shape[0], nclass, device=dev); loss = torch.zeros(1, device=dev)
            out, _ = eng.run(g, eng.pack(xf.to(dev), xr.to(dev)), t.to(dev), probs, loss, train=False)
            with torch.no_grad():
                _, a, _, _ = m(xf.to(dev), g, None)
                _, b, _, _ = m(xr.to(dev), g, None)
            print(which, "impl", impl, c, "n=%d" % xf.shape[0], "engine probs err %.2e" % ogcn.max_rel(probs.cpu(), want),
                  "engine logits err %.2e" % ogcn.max_rel(out.mean(1).cpu(), (pf + pr) / 2),
                  "module err %.2e" % ogcn.max_rel(((a + b) / 2).cpu(), (pf + pr) / 2))
