import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chromegcn_b200 import ops
dev = torch.device("cuda", 0)
m = int(sys.argv[1]) if len(sys.argv) > 1 else 113000
a = torch.randn(m, 128, device=dev); w = torch.randn(128, 128, device=dev); bias = torch.randn(128, device=dev)
for _ in range(3):
    ops.gemm_rowpanel(a, w, False, bias, impl=2)
torch.cuda.synchronize()
