import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chromegcn_b200 import ops
dev = torch.device("cuda", 0)
torch.manual_seed(0)
sizes = [1, 7, 31, 32, 33, 58, 64, 80, 127, 128, 129, 255, 256, 257, 314, 422, 511, 513, 1000, 4097, 20001]
bad = 0
for rep in range(2):
    for m in sizes:
        a = torch.randn(m, 128, device=dev); b = torch.randn(m, 128, device=dev); w = torch.randn(128, 128, device=dev) * 0.1
        want = a.double() @ w.double()
        got = ops.gemm_rowpanel(a, w, False, None, impl=2)
        e1 = float((got.double() - want).abs().max() / want.abs().max())
        wantg = a.double().t() @ b.double()
        gotg = ops.gemm_gram(a, b, impl=2)
        e2 = float((gotg.double() - wantg).abs().max() / wantg.abs().max())
        flag = "" if (e1 < 1e-5 and e2 < 1e-5) else "   <<<<<< BAD"
        bad += bool(flag)
        print("m=%6d rowpanel %.2e gram %.2e%s" % (m, e1, e2, flag))
print("bad:", bad)
