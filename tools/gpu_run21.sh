cd $GRAFT_REPO_ROOT
timeout -k 10 300 python __graft_entry__.py --smoke 2>&1 | grep -E "smoke|Error|error" | head -5
timeout -k 10 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -12
( time timeout -k 10 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02_bench_wg_1gpu.json 2> gpurun_out/bench21_err.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_wg_1gpu.json | head -2; tail -4 gpurun_out/bench21_err.log
