"""Text ingest of one Juicer RAWobserved dump: native parser (cgcn_contacts_parse) vs pandas C parser vs the
reference's csv.DictReader loop (data/7create_graph_new.py:71-76, bounded sample).  Host only."""
import csv, json, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from chromegcn_b200 import create_graph as cg

rows = int(float(sys.argv[1])) if len(sys.argv) > 1 else 20_000_000
rng = np.random.default_rng(0)
b1 = rng.integers(0, 250_000, rows) * 1000
b2 = b1 + rng.integers(0, 3000, rows) * 1000
v = np.floor(rng.random(rows) * 300) + 1
path = os.path.join(tempfile.mkdtemp(), "chr1_1kb.RAWobserved")
import pandas as pd
pd.DataFrame({"a": b1, "b": b2, "v": v}).to_csv(path, sep="\t", header=False, index=False)
size = os.path.getsize(path)
t0 = time.perf_counter(); a = cg.read_contacts(path); t_native = time.perf_counter() - t0
t0 = time.perf_counter(); a = cg.read_contacts(path); t_native = min(t_native, time.perf_counter() - t0)
os.environ["CGCN_TEXT_PARSER"] = "pandas"
t0 = time.perf_counter(); b = cg.read_contacts(path); t_pandas = time.perf_counter() - t0
assert all(np.array_equal(x, y) for x, y in zip(a, b))
sample = min(rows, 1_000_000)
t0 = time.perf_counter()
with open(path) as fp:
    d = {}
    for i, line in enumerate(csv.DictReader(fp, delimiter="\t", fieldnames=["start_pos1", "start_pos2", "val"])):
        if i >= sample:
            break
        d[(int(line["start_pos1"]), int(line["start_pos2"]))] = float(line["val"])
t_ref = (time.perf_counter() - t0) * rows / sample
print(json.dumps({"rows": rows, "file_MB": size / 1e6, "cores": os.cpu_count(), "native_s": t_native, "native_MBps": size / 1e6 / t_native,
                  "pandas_s": t_pandas, "reference_csv_loop_s_extrapolated": t_ref, "speedup_vs_pandas": t_pandas / t_native,
                  "speedup_vs_reference_loop": t_ref / t_native}))
