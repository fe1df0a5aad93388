"""Quick GPU check of the tcgen05 contractions against fp64 and the FFMA kernels (run under `timeout`)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chromegcn_b200 import ops

dev = torch.device("cuda", 0)
torch.manual_seed(0)


def timeit(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "rowpanel"):
    for m in (128, 1000, 40000, 102000, 194330, 400000):
        for bt in (False, True):
            a = torch.randn(m, 128, device=dev)
            w = torch.randn(128, 128, device=dev) * 0.1
            bias = torch.randn(128, device=dev)
            want = a.double() @ (w.double().t() if bt else w.double()) + bias.double()
            got2 = ops.gemm_rowpanel(a, w, bt, bias, impl=2)
            got1 = ops.gemm_rowpanel(a, w, bt, bias, impl=1)
            torch.cuda.synchronize()
            e2 = float((got2.double() - want).abs().max() / want.abs().max())
            e1 = float((got1.double() - want).abs().max() / want.abs().max())
            t2 = timeit(lambda: ops.gemm_rowpanel(a, w, bt, bias, impl=2))
            t1 = timeit(lambda: ops.gemm_rowpanel(a, w, bt, bias, impl=1))
            print("rowpanel m=%6d bt=%d  tc err %.2e (%.1f us)   ffma err %.2e (%.1f us)   GB/s tc %.0f" %
                  (m, bt, e2, t2, e1, t1, 2 * m * 512 / t2 / 1e3), flush=True)
if which in ("all", "gram"):
    for m in (32, 1000, 40000, 102000, 194330, 400000):
        a = torch.randn(m, 128, device=dev)
        b = torch.randn(m, 128, device=dev)
        want = a.double().t() @ b.double()
        got2 = ops.gemm_gram(a, b, impl=2)
        got1 = ops.gemm_gram(a, b, impl=1)
        torch.cuda.synchronize()
        e2 = float((got2.double() - want).abs().max() / want.abs().max())
        e1 = float((got1.double() - want).abs().max() / want.abs().max())
        t2 = timeit(lambda: ops.gemm_gram(a, b, impl=2))
        t1 = timeit(lambda: ops.gemm_gram(a, b, impl=1))
        print("gram     m=%6d       tc err %.2e (%.1f us)   ffma err %.2e (%.1f us)   GB/s tc %.0f" %
              (m, e2, t2, e1, t1, 2 * m * 512 / t2 / 1e3), flush=True)
