#!/bin/bash
# gram look-ahead A/B, locality sensitivity lines (1 GPU)
mkdir -p gpurun_out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e"
timeout 300 python -m pytest tests/test_gpu_kernels.py -q -x -k "gram or rowpanel" > gpurun_out/run25_pytest.log 2>&1; tail -2 gpurun_out/run25_pytest.log
CGCN_GRAM_SLOTS=2 timeout 300 python bench.py $B --no-roofline > gpurun_out/run25_slots2.json 2> gpurun_out/run25_slots2.err
timeout 300 python bench.py $B --no-roofline > gpurun_out/run25_slots3.json 2> gpurun_out/run25_slots3.err
timeout 400 python bench.py $B --graph permuted > gpurun_out/r02_bench_wg_1gpu_permuted.json 2> gpurun_out/run25_perm.err
timeout 400 python bench.py $B --graph longrange > gpurun_out/r02_bench_wg_1gpu_longrange.json 2> gpurun_out/run25_long.err
for f in run25_slots2 run25_slots3 r02_bench_wg_1gpu_permuted r02_bench_wg_1gpu_longrange; do python - <<PY
import json
try:
    d=json.loads([l for l in open('gpurun_out/$f.json') if l.startswith('{')][-1])
    r=d.get('roofline') or {}
    print('$f', round(d['ms_per_step'],3), 'ms', round(d['value'],3), 'GE/s', 'fused_us', r.get('avg_launch_us'), 'tbf', r.get('time_bound_frac'), 'spmm', (r.get('spmm') or {}).get('avg_launch_us'))
except Exception as e:
    print('$f', 'FAILED', e)
PY
done
