#!/bin/bash
# 8 epilogue warps with setmaxnreg (gather 96 regs, epilogue 56) vs the default 4 epilogue warps
mkdir -p gpurun_out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-roofline"
timeout 400 python -m pytest tests/test_gpu_model.py -q -x -k "alternative_kernel_paths" > gpurun_out/run26_pytest.log 2>&1; tail -3 gpurun_out/run26_pytest.log
timeout 300 python bench.py $B > gpurun_out/run26_e4.json 2> gpurun_out/run26_e4.err
CGCN_FUSED_EPI=8 timeout 300 python bench.py $B > gpurun_out/run26_e8.json 2> gpurun_out/run26_e8.err
CGCN_FUSED_EPI=8 timeout 300 python bench.py $B > gpurun_out/run26_e8b.json 2> gpurun_out/run26_e8b.err
for f in run26_e4 run26_e8 run26_e8b; do grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.json | head -1; done
