"""CPU, build container only (skipped where /root/reference does not exist, e.g. on the GPU box): the oracle against the
LIVE reference on seeds other than the ones the committed golden vectors were minted with -- the adjacency build
(data/7create_graph_new.py create_graph, unmodified), process_graph('hic') (utils/util_methods.py) and the fp32 model
forward + autograd (models/ChromeModels.py).  Bit-identical in all three."""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

REF = os.environ.get("CHROMEGCN_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models")), reason="live reference not present")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def mg():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)              # puts the reference on sys.path exactly like the minting run
    return mod


@pytest.mark.parametrize("seed,norm,hic_edges", [(501, "SQRTVC", 900), (502, "", 700), (503, "SQRTVC", 100000)])
def test_adjacency_oracles_equal_live_create_graph(mg, seed, norm, hic_edges):
    from oracle import adjacency as oadj
    hics = [mg.adversarial_hic(c, 400, 140 + 9 * i, 2500, seed + i) for i, c in enumerate(["chr1", "chr2", "chr3", "chr22"])]
    if norm == "":                             # the un-normalised mode reads a file pre-sorted by value, descending
        for h in hics:
            order = np.argsort(-h.val, kind="stable")
            h.bin1, h.bin2, h.val = h.bin1[order], h.bin2[order], h.val[order]
    graphs = mg.run_reference_create_graph(hics, norm, hic_edges)
    for h in hics:
        csr = graphs[h.chrom]
        for build in (oadj.build_adjacency_loops, oadj.build_adjacency_numpy):
            ip, ix = build(h.window_starts, h.bin1, h.bin2, h.val, h.norm if norm else None, 1, hic_edges)
            assert np.array_equal(ip, csr.indptr) and np.array_equal(ix, csr.indices), (h.chrom, build.__name__)


def test_process_graph_and_model_equal_live_reference(mg):
    from oracle import adjacency as oadj
    from oracle import gcn as ogcn
    from models.ChromeModels import ChromeGCN as RefGCN                      # the reference's own class
    from utils import util_methods as ref_util
    from scipy import sparse
    h = mg.adversarial_hic("chr5", 500, 230, 4000, 777)
    ip, ix = oadj.build_adjacency_numpy(h.window_starts, h.bin1, h.bin2, h.val, h.norm, 1, 1500)
    n = ip.shape[0] - 1
    csr = sparse.csr_matrix((np.ones(ix.shape[0]), ix, ip), shape=(n, n))
    ref_adj = ref_util.process_graph("hic", {"chr5": csr}, n, "chr5")
    r, c, val = oadj.normalize_hic(ip, ix, n)
    assert np.array_equal(ref_adj._indices().numpy(), np.stack([r, c])) and np.array_equal(ref_adj._values().numpy(), val)
    torch.manual_seed(11)
    ref = RefGCN(128, 128, 13, 0.0, True, 2)
    mine = ogcn.ChromeGCNOracle(128, 128, 13, 0.0, True, 2)
    mine.load_state_dict(ref.state_dict())
    ogcn.stress_init_(ref)
    mine.load_state_dict(ref.state_dict())
    gen = torch.Generator().manual_seed(12)
    x = torch.randn(n, 128, generator=gen)
    t = (torch.rand(n, 13, generator=gen) < 0.2).float()
    torch.set_num_threads(1)
    outs = []
    for model in (ref, mine):
        model.train()
        xi = x.clone().requires_grad_(True)
        _, out, (g1, g2), _ = model(xi, ref_adj, None)
        loss = torch.nn.functional.binary_cross_entropy_with_logits(out, t)
        loss.backward()
        outs.append([out.detach(), g1.detach(), g2.detach(), xi.grad] + [p.grad for p in model.parameters()])
    for a, b in zip(*outs):
        assert torch.equal(a, b)
