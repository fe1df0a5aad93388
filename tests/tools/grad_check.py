"""Per-parameter gradient error of the fused step on a chr22-sized graph, FFMA vs tcgen05 contractions."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from chromegcn_b200 import synthetic
from chromegcn_b200.chrome_models import ChromeGCN
from chromegcn_b200.engine import ChromosomeEngine
from chromegcn_b200.graph import HiCGraph
from oracle import adjacency as oadj, gcn as ogcn
dev = torch.device("cuda", 0)
chrom = sys.argv[1] if len(sys.argv) > 1 else "chr22"
h = synthetic.make_hic(chrom)
ip, ix = oadj.build_adjacency_numpy(h.window_starts, h.bin1, h.bin2, h.val, h.norm, 1, 500000)
n = ip.shape[0] - 1
feats = synthetic.make_features(chrom, n)
for init in ("stress", "reference"):
    torch.manual_seed(0)
    om = ogcn.ChromeGCNOracle(128, 128, synthetic.NCLASS, 0.0, True, 2)
    if init == "stress":
        ogcn.stress_init_(om)
    sd = {k: v.clone() for k, v in om.state_dict().items()}
    o32 = ogcn.ChromeGCNOracle(128, 128, synthetic.NCLASS, 0.0, True, 2); o32.load_state_dict(sd); o32.train()
    om = om.double().train()
    adj = ogcn.coo_adjacency(ip, ix, torch.float64)
    lo, _, pred_o, _ = ogcn.chromosome_step(om, feats["forward"].double(), feats["backward"].double(), feats["target"].double(), adj, None, True)
    ogcn.chromosome_step(o32, feats["forward"], feats["backward"], feats["target"], ogcn.coo_adjacency(ip, ix), None, True)
    g = HiCGraph.from_csr_pattern(ip, ix, dev)
    res = {}
    for impl in (1, 0):
        m = ChromeGCN(128, 128, synthetic.NCLASS, 0.0, True, 2); m.load_state_dict(sd); m = m.to(dev).train(); m.gemm_impl = impl
        eng = ChromosomeEngine(m, 2)
        loss = torch.zeros(1, device=dev)
        out, _ = eng.run(g, eng.pack(feats["forward"].to(dev), feats["backward"].to(dev)), feats["target"].to(dev), None, loss, train=True)
        res[impl] = ({k: p.grad.detach().cpu().clone() for k, p in m.named_parameters()}, ogcn.max_rel(out.mean(1).cpu(), pred_o))
    print("init=%s n=%d  pred err: ffma %.2e  tc %.2e" % (init, n, res[1][1], res[0][1]))
    for (k, q), (_, q32) in zip(om.named_parameters(), o32.named_parameters()):
        print("  %-20s  torch-fp32 %.2e   ffma %.2e   tc %.2e" % (k, ogcn.max_rel(q32.grad, q.grad), ogcn.max_rel(res[1][0][k], q.grad), ogcn.max_rel(res[0][0][k], q.grad)))
