"""CPU oracle for the ChromeGCN chromosome-model path.

TEST INFRASTRUCTURE ONLY.  This package is a clean-room CPU restatement of the
reference algorithms (QData/ChromeGCN) that the CUDA path in `chromegcn_b200/` is
checked against.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it -- as the checker or as the
timed CPU baseline, never as the product.  `chromegcn_b200/` must not import it.

Parity pin: the reference ships no tests, golden vectors or known-answer fixtures
for this path, so the pin is the reference itself,
executed in the build container: `tests/golden/make_golden.py` imports
`/root/reference` (models/ChromeModels.py, utils/util_methods.py, finetune.py,
data/7create_graph_new.py), runs it on seeded inputs and commits inputs + outputs
as `tests/golden/*.npz`.  `tests/test_oracle_golden.py` checks every function here
against those vectors (bit-exact for the integer work, <= 2e-6 for fp32).

Modules
  adjacency.py  Hi-C contacts -> binary symmetric CSR   (data/7create_graph_new.py:14-120)
                + D^-1 (A+I) normalisation -> COO       (utils/util_methods.py:99-106,120-135,146-180)
  gcn.py        GraphConvolution / ChromeGCN forward, the finetune train step and the
                optimiser (models/SubLayers.py:42-52, models/ChromeModels.py:22-52,
                finetune.py:29-53, utils/util_methods.py:14-19)
"""
