"""CPU: the memory-mapped multi-threaded text ingest (cgcn_contacts_parse / cgcn_vector_parse, csrc/ingest.cu) against
Python's own int() / float() on the same tokens -- what csv.DictReader + int()/float() produce in
data/7create_graph_new.py:56-59,71-76 -- and against the pandas route of create_graph.py."""
import os

import numpy as np
import pytest

from chromegcn_b200 import _lib, create_graph as cg


def _write_contacts(path, n, seed, crlf=False, trailing_newline=True):
    rng = np.random.default_rng(seed)
    b1 = rng.integers(0, 250_000, n) * 1000
    b2 = b1 + rng.integers(0, 3000, n) * 1000
    kind = rng.integers(0, 6, n)
    vals = []
    for k, r in zip(kind, rng.random(n)):
        if k == 0:
            vals.append("%d.0" % int(1 + r * 500))                  # Juicer's usual "12.0"
        elif k == 1:
            vals.append(repr(float(r * 1e3)))                       # shortest round-trip repr
        elif k == 2:
            vals.append("%.17g" % (r * 1e-3))                       # 17 significant digits
        elif k == 3:
            vals.append("%e" % (r * 1e-5))                          # exponent form
        elif k == 4:
            vals.append(str(int(r * 90)))                           # bare integer
        else:
            vals.append("%.3f" % (r * 40))
    eol = "\r\n" if crlf else "\n"
    lines = ["%d\t%d\t%s" % t for t in zip(b1, b2, vals)]
    text = eol.join(lines) + (eol if trailing_newline else "")
    with open(path, "w", newline="") as fp:
        fp.write(text)
    return b1, b2, np.array([float(v) for v in vals], dtype=np.float64)


@pytest.mark.parametrize("n,crlf,trailing,threads", [(300_000, False, True, 0), (300_000, True, False, 7), (5, False, False, 3),
                                                    (1, False, True, 1)])
def test_contacts_parse_equals_python_float(tmp_path, n, crlf, trailing, threads):
    path = str(tmp_path / "chrT_1kb.RAWobserved")
    b1, b2, v = _write_contacts(path, n, n + threads, crlf, trailing)
    g1, g2, gv = cg.read_contacts(path, threads=threads)
    assert g1.dtype == np.int64 and gv.dtype == np.float64 and len(g1) == n
    assert np.array_equal(g1, b1) and np.array_equal(g2, b2)
    assert np.array_equal(gv.view(np.uint64), v.view(np.uint64))               # bit-identical doubles


def test_native_and_pandas_routes_agree(tmp_path, monkeypatch):
    path = str(tmp_path / "c.RAWobserved")
    _write_contacts(path, 50_000, 11)
    a = cg.read_contacts(path)
    monkeypatch.setenv("CGCN_TEXT_PARSER", "pandas")
    b = cg.read_contacts(path)
    for x, y in zip(a, b):
        assert x.dtype == y.dtype and np.array_equal(x.view(np.uint64) if x.dtype == np.float64 else x,
                                                     y.view(np.uint64) if y.dtype == np.float64 else y)


def test_vector_parse_nan_blank_lines_and_extra_columns(tmp_path, monkeypatch):
    path = str(tmp_path / "chrT_1kb.SQRTVCnorm")
    toks = ["1.0270898", "NaN", "0.0", "nan", "2", "1e-3", "+1.5", " 3.25 ", "inf", "0.30000000000000004"]
    with open(path, "w") as fp:
        fp.write("\n".join(toks[:5]) + "\n\n   \n" + "\n".join(toks[5:]) + "\n")
    got = cg.read_norm_vector(path)
    want = np.array([float(t) for t in toks])
    assert len(got) == len(want)
    assert np.array_equal(np.isnan(got), np.isnan(want))
    ok = ~np.isnan(want)
    assert np.array_equal(got[ok].view(np.uint64), want[ok].view(np.uint64))
    # further tab-separated columns on a contact row are ignored (csv.DictReader puts them under restkey)
    p2 = str(tmp_path / "extra.RAWobserved")
    with open(p2, "w") as fp:
        fp.write("1000\t2000\t3.5\tignored\n5000\t5000\t1\n")
    b1, b2, v = cg.read_contacts(p2)
    assert b1.tolist() == [1000, 5000] and b2.tolist() == [2000, 5000] and v.tolist() == [3.5, 1.0]


def test_malformed_rows_are_reported_with_their_row(tmp_path):
    path = str(tmp_path / "bad.RAWobserved")
    with open(path, "w") as fp:
        fp.write("1000\t2000\t3.5\n1000\t2000.0\t1.0\n")           # int('2000.0') raises in the reference too
    with pytest.raises(_lib.ChromeGCNNativeError, match="row 1"):
        cg.read_contacts(path)
    with open(path, "w") as fp:
        fp.write("1000\t2000\n")
    with pytest.raises(_lib.ChromeGCNNativeError, match="three tab-separated"):
        cg.read_contacts(path)
    with pytest.raises(_lib.ChromeGCNNativeError, match="cannot open"):
        cg.read_contacts(str(tmp_path / "missing"))
    empty = str(tmp_path / "empty")
    open(empty, "w").close()
    b1, b2, v = cg.read_contacts(empty)
    assert len(b1) == 0 and len(v) == 0


def test_bed_window_starts_match_the_reference_loop(tmp_path, monkeypatch):
    """create_bin_dict (data/7create_graph_new.py:24-44): unique start positions per requested chromosome, sorted."""
    rng = np.random.default_rng(3)
    chroms_in_file = ["chr1", "chr10", "chr2", "chrX", "chr1_gl000191_random"]
    rows = []
    for _ in range(60_000):
        c = chroms_in_file[rng.integers(0, len(chroms_in_file))]
        s = int(rng.integers(0, 3000)) * 1000
        rows.append("%s\t%d\t%d\tassay%d\t0\t.\t1.5\t2.0\t3.0\t50" % (c, s, s + 1000, rng.integers(0, 100)))
    path = str(tmp_path / "chipseq_windows.bed")
    with open(path, "w") as fp:
        fp.write("\n".join(rows) + "\n")
    wanted = ["chr1", "chr2", "chr3", "chrX"]                      # chr3 has no rows; chr10 / the random contig are ignored
    want = {c: set() for c in wanted}
    for r in rows:
        f = r.split("\t")
        if f[0] in want:
            want[f[0]].add(int(f[1]))
    got = cg.read_window_starts(path, wanted)
    monkeypatch.setenv("CGCN_TEXT_PARSER", "pandas")
    got_pd = cg.read_window_starts(path, wanted)
    for c in wanted:
        assert got[c].dtype == np.int64 and got[c].tolist() == sorted(want[c])
        assert np.array_equal(got[c], got_pd[c])
