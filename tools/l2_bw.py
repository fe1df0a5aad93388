"""L2 and HBM read bandwidth of this GPU, measured with libchromegcn's probe kernel (cgcn_membw_read): the denominators of
the time-bound roofline bench.py reports for the gather kernels (an SpMM whose panel is L2 resident is bounded by the
L2 -> SM rate on its ALGORITHMIC bytes and by HBM on its compulsory bytes).  Prints one JSON line; the committed copy is
profiles/r02_l2_bw.json."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from chromegcn_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda", 0)
sink = torch.zeros(1, device=dev)
out = {}
for name, mb, reps in (("l2_16MB", 16, 400), ("l2_32MB", 32, 200), ("l2_64MB", 64, 100), ("hbm_4GB", 4096, 2)):
    buf = torch.randn(mb * 1024 * 1024 // 4, device=dev)
    for _ in range(2):
        _lib.check(lib.cgcn_membw_read(buf.data_ptr(), buf.numel() * 4, reps, sink.data_ptr(), _lib.current_stream()))
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(lib.cgcn_membw_read(buf.data_ptr(), buf.numel() * 4, reps, sink.data_ptr(), _lib.current_stream()))
        e1.record()
        e1.synchronize()
        best = max(best, buf.numel() * 4.0 * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9)
    out[name + "_GBs"] = round(best, 1)
    del buf
out["l2_read_peak_GBs"] = max(out["l2_16MB_GBs"], out["l2_32MB_GBs"], out["l2_64MB_GBs"])
out["how"] = "cgcn_membw_read: grid = 8 CTAs per SM x 256 threads, 8 independent ld.global.cg.v4 per thread, best of 5"
out["gpu"] = torch.cuda.get_device_name(0)
print(json.dumps(out))
