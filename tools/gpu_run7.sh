cd $GRAFT_REPO_ROOT
timeout -k 10 300 python __graft_entry__.py --smoke 2>&1 | grep -E "smoke|Error|error" | head -5
timeout -k 10 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py -m gpu -q --timeout 180 -x 2>&1 | tail -2
for cfg in "CGCN_FUSED_GW=16" "CGCN_FUSED_GW=8"; do
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench10_$tag.log 2>&1
  echo "$cfg: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench10_$tag.log | head -1) $(grep -o '"final_loss_sum": [0-9.]*' gpurun_out/r02_bench10_$tag.log)"
done
ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_fused10.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_layer -s 48 -c 3 -f -o gpurun_out/r02_fused10 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b.log 2>&1
