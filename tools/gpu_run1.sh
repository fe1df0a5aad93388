set -x
cd $GRAFT_REPO_ROOT
timeout -k 10 300 python __graft_entry__.py --smoke > gpurun_out/r02_smoke1.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke1.log
tail -5 gpurun_out/r02_smoke1.log
timeout -k 10 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py tests/test_gpu_rowpartition.py -m gpu -q --timeout 180 -x 2>&1 | tail -40 > gpurun_out/r02_pytest1.log
tail -40 gpurun_out/r02_pytest1.log
for cfg in "CGCN_NO_FUSED=1" "CGCN_FUSED_GW=16" "CGCN_FUSED_GW=8"; do
  env $cfg timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench1_$cfg.log 2>&1
  echo "$cfg: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench1_$cfg.log | head -1) $(grep -o '"final_loss_sum": [0-9.]*' gpurun_out/r02_bench1_$cfg.log)"
done
