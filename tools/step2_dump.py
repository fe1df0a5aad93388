import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from chromegcn_b200.chrome_models import ChromeGCN
from chromegcn_b200.engine import ChromosomeEngine
from chromegcn_b200.graph import HiCGraph
from chromegcn_b200.optim import FlatSGD
dev = torch.device("cuda", 0)
z = np.load("tests/golden/finetune.npz")
nclass = int(z["nclass"])
sd = {k[4:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd0.")}
def data(c):
    ip, ix = z[c + ".indptr"], z[c + ".indices"]
    return ip, ix, [torch.from_numpy(z["%s.%s" % (c, k)]) for k in ("forward", "backward", "target")]
def layout(n, S=2, d=128):
    off = 0; L = {}
    def take(name, fl):
        nonlocal off
        off = (off + 63) // 64 * 64; L[name] = (off, fl); off += fl
    panel = n * S * d
    for l in range(2):
        take("ax%d" % l, panel); take("z%d" % l, panel); take("xo%d" % l, panel)
    take("hb", panel)
    for k in ("bn_mean", "bn_rstd", "bn_c1", "bn_c2"): take(k, S * d)
    for k in ("dA", "dB", "dC"): take(k, panel)
    return L
# bring a model to the post-step-1 state with FFMA, then evaluate step 2 with both impls from identical weights
m0 = ChromeGCN(128, 128, nclass, 0.0, True, 2); m0.load_state_dict(sd); m0 = m0.to(dev).train(); m0.gemm_impl = 1
eng = ChromosomeEngine(m0, 2); opt = FlatSGD(m0, lr=0.25)
ip, ix, (xf, xr, t) = data("chr1")
eng.run(HiCGraph.from_csr_pattern(ip, ix, dev), eng.pack(xf.to(dev), xr.to(dev)), t.to(dev), None, torch.zeros(1, device=dev), train=True)
opt.step()
state = {k: v.detach().clone() for k, v in m0.state_dict().items()}
ip, ix, (xf, xr, t) = data("chr2")
n = xf.shape[0]
g = HiCGraph.from_csr_pattern(ip, ix, dev)
dumps = {}
for impl in (1, 0):
    m = ChromeGCN(128, 128, nclass, 0.0, True, 2); m.load_state_dict(state); m = m.to(dev).train(); m.gemm_impl = impl
    e = ChromosomeEngine(m, 2)
    e.run(g, e.pack(xf.to(dev), xr.to(dev)), t.to(dev), None, torch.zeros(1, device=dev), train=True)
    torch.cuda.synchronize()
    ws = e._bufs["ws"]
    d = {k: ws[o:o + fl].clone() for k, (o, fl) in layout(n).items()}
    for bk in ("gate0", "gate1", "out", "dout", "panel"):
        d["buf." + bk] = e._bufs[bk][: (n * 2 * (128 if bk == "panel" else (40 if bk in ("out", "dout") else 1)))].clone()
    dumps[impl] = (d, {k: p.grad.clone() for k, p in m.named_parameters()})
for k in dumps[1][0]:
    a, b = dumps[1][0][k], dumps[0][0][k]
    print("%-8s ffma-vs-tc max abs diff %.3e  (max |ffma| %.3e)" % (k, float((a - b).abs().max()), float(a.abs().max())))
for k in dumps[1][1]:
    a, b = dumps[1][1][k], dumps[0][1][k]
    print("grad %-20s diff %.3e (max %.3e)" % (k, float((a - b).abs().max()), float(a.abs().max())))
w = state["GC1.weight"]
print("GC1.weight finite:", bool(torch.isfinite(w).all()), "min|w| %.3e" % float(w.abs().min()), "bias GC1 max %.3e" % float(state["GC1.bias"].abs().max()))

for k in ("dC", "dB", "dA", "z1", "xo1", "hb"):
    a, b = dumps[1][0][k].view(n, 2, 128), dumps[0][0][k].view(n, 2, 128)
    diff = (a - b).abs()
    rows = diff.amax(dim=(1, 2)); cols = diff.amax(dim=(0, 1)); st = diff.amax(dim=(0, 2))
    thr = float(diff.max()) * 0.1
    bad_rows = torch.nonzero(rows > thr).flatten().tolist()
    bad_cols = torch.nonzero(cols > thr).flatten().tolist()
    print(k, "max", float(diff.max()), "bad rows (n=%d):" % len(bad_rows), bad_rows[:12], "... bad cols (n=%d):" % len(bad_cols), bad_cols[:12], "strand max", st.tolist())
