#!/bin/bash
# final 1-GPU verification: whole GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout -k 5 330 python -m pytest tests -m gpu -x -q > gpurun_out/r02_pytest_1gpu_final.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_pytest_1gpu_final.log
timeout -k 5 90 python __graft_entry__.py --smoke > gpurun_out/r02_smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r02_smoke_final.log
timeout -k 5 200 python bench.py > gpurun_out/r02_bench_wg_1gpu_final.json 2> gpurun_out/r02_bench_final.err; echo "bench rc=$?"
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_wg_1gpu_final.json | head -2
