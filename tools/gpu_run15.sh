cd $GRAFT_REPO_ROOT
timeout -k 10 300 python __graft_entry__.py --smoke 2>&1 | grep -E "smoke|Error|error" | head -5
timeout -k 10 1200 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py tests/test_gpu_rowpartition.py -m gpu -q --timeout 400 2>&1 | tail -4
for cfg in "CGCN_FUSED_EPI=8 CGCN_LAYER_MODE=gather" "CGCN_FUSED_EPI=4 CGCN_LAYER_MODE=gather" "CGCN_FUSED_EPI=8 CGCN_LAYER_MODE=stream CGCN_FUSED_HEAD=1" "CGCN_FUSED_EPI=8 CGCN_LAYER_MODE=gather CGCN_FUSED_HEAD=1"; do
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-roofline > gpurun_out/r02_bench15_$tag.log 2>&1
  echo "$cfg: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench15_$tag.log | head -1) $(grep -o '"final_loss_sum": [0-9.]*' gpurun_out/r02_bench15_$tag.log)"
done
CGCN_FUSED_HEAD=1 ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_15.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-roofline > gpurun_out/ncu_a.log 2>&1
