#!/bin/bash
# 8 epilogue warps, setmaxnreg 64 / 88 (increments covered by the CTA's own decrements), hand-over before barrier 3
mkdir -p gpurun_out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-roofline"
CGCN_FUSED_EPI=8 timeout -k 5 60 python bench.py $B > gpurun_out/run31_e8.json 2> gpurun_out/run31_e8.err || { echo "e8 run failed or timed out"; exit 0; }
grep -o '"ms_per_step": [0-9.]*' gpurun_out/run31_e8.json | head -1; grep -o '"final_loss_sum": [0-9.e+-]*' gpurun_out/run31_e8.json | head -1
timeout -k 5 80 python -m pytest tests/test_gpu_model.py -q -x -k "alternative_kernel_paths" > gpurun_out/run31_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/run31_pytest.log
