"""CPU property tests (hypothesis) for the host logic and the oracle: the two restatements of the adjacency contract
agree on arbitrary small inputs (ties, duplicate and reversed keys, NaN / 0 norm entries, non-window rows, diagonal
rows, K larger and smaller than the candidate set, the un-normalised early-exit mode), and the result has the
structural properties data/7create_graph_new.py:108-120 guarantees; the label bit rows round-trip; the chromosome
schedule covers every chromosome exactly once for every world size."""
import numpy as np
import torch
from hypothesis import given, settings, strategies as st

from chromegcn_b200 import dist as cdist
from chromegcn_b200 import ops
from oracle import adjacency as oadj


@st.composite
def hic_case(draw):
    n_bins = draw(st.integers(4, 40))
    starts = sorted(draw(st.sets(st.integers(0, n_bins - 1), min_size=1, max_size=n_bins)))
    m = draw(st.integers(0, 120))
    b1 = draw(st.lists(st.integers(0, n_bins - 1), min_size=m, max_size=m))
    b2 = draw(st.lists(st.integers(0, n_bins - 1), min_size=m, max_size=m))
    vals = draw(st.lists(st.sampled_from([0.0, 1.0, 2.0, 2.0, 3.0, 5.5, 7.0]), min_size=m, max_size=m))   # heavy ties
    norm = draw(st.lists(st.sampled_from([1.0, 0.5, 2.0, float("nan"), 0.0, 1.25]), min_size=n_bins, max_size=n_bins))
    use_norm = draw(st.booleans())
    hic_edges = draw(st.integers(0, 2 * m + 4))
    return (np.array(starts, dtype=np.int64) * 1000, np.array(b1, dtype=np.int64) * 1000, np.array(b2, dtype=np.int64) * 1000,
            np.array(vals, dtype=np.float64), np.array(norm, dtype=np.float64) if use_norm else None, hic_edges)


@settings(max_examples=150, deadline=None)
@given(hic_case())
def test_adjacency_restatements_agree_and_are_well_formed(case):
    starts, b1, b2, v, norm, hic_edges = case
    ip1, ix1 = oadj.build_adjacency_loops(starts, b1, b2, v, norm, 1, hic_edges)
    ip2, ix2 = oadj.build_adjacency_numpy(starts, b1, b2, v, norm, 1, hic_edges)
    assert np.array_equal(ip1, ip2) and np.array_equal(ix1, ix2)
    n = len(starts)
    assert ip1.shape == (n + 1,) and ip1[0] == 0 and ip1[-1] == ix1.shape[0]
    rows = np.repeat(np.arange(n), np.diff(ip1))
    assert not np.any(rows == ix1)                                            # no diagonal (:78 drops bin1 == bin2)
    pairs = set(zip(rows.tolist(), ix1.tolist()))
    assert all((j, i) in pairs for i, j in pairs)                            # symmetric (:115-116)
    assert len(pairs) == ix1.shape[0]                                         # no duplicate entries
    assert all(np.all(np.diff(ix1[ip1[r]: ip1[r + 1]]) > 0) for r in range(n))  # ascending columns
    k_pairs = int(hic_edges / 2.0)
    if k_pairs > 0:                                                           # K == 0 keeps everything (the loop at
        assert ix1.shape[0] <= 2 * k_pairs                                    # :98-102 never sees idx == 0)
    # the online normalisation on top: D^-1 (A + I) has row sums 1 and 1/deg values (utils/util_methods.py:99-106)
    rp, ci = oadj.pattern_with_selfloops(ip1, ix1)
    assert np.all(np.diff(rp) == np.diff(ip1) + 1)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 70), st.integers(1, 200), st.integers(0, 2 ** 31 - 1))
def test_label_bit_rows_roundtrip(n, c, seed):
    rng = np.random.default_rng(seed)
    t = (rng.random((n, c)) < 0.5).astype(np.float32)
    b = ops.pack_targets(torch.from_numpy(t))
    w = b.numpy().view(np.uint32)
    cols = np.arange(c)
    back = ((w[:, cols >> 5] >> (cols & 31).astype(np.uint32)) & 1).astype(np.float32)
    assert np.array_equal(back, t)


@settings(max_examples=60, deadline=None)
@given(st.lists(st.floats(1.0, 100.0), min_size=1, max_size=30), st.integers(1, 8))
def test_schedule_covers_every_chromosome_once(costs, world):
    names = {"c%02d" % i: c for i, c in enumerate(costs)}
    sch = cdist.balanced_schedule(names, world)
    assert all(len(r) == world for r in sch)
    flat = [c for r in sch for cell in r for c in cell]
    assert sorted(flat) == sorted(names)
    assert sorted(sum(cdist.schedule_shards(sch, world), [])) == sorted(names)
    assert cdist.schedule_cost(sch, names) >= sum(costs) / world - 1e-9       # lock-step lower bound
