#!/bin/bash
# 8 epilogue warps, setmaxnreg 56 / 96, register hand-over before barrier 3
mkdir -p gpurun_out
B="--steps 20 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-roofline"
CGCN_FUSED_EPI=8 timeout -k 5 70 python bench.py $B > gpurun_out/run30_e8.json 2> gpurun_out/run30_e8.err || { echo "e8 run failed or timed out"; exit 0; }
CGCN_FUSED_EPI=8 timeout -k 5 70 python bench.py $B > gpurun_out/run30_e8b.json 2> gpurun_out/run30_e8b.err
for f in run30_e8 run30_e8b; do grep -o '"ms_per_step": [0-9.]*' gpurun_out/$f.json | head -1; grep -o '"final_loss_sum": [0-9.e+-]*' gpurun_out/$f.json | head -1; done
