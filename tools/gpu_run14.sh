cd $GRAFT_REPO_ROOT
timeout -k 10 300 python __graft_entry__.py --smoke 2>&1 | grep -E "smoke|Error|error" | head -5
timeout -k 10 1200 python -m pytest tests -m gpu -q --timeout 400 2>&1 | tail -8
for cfg in "CGCN_LAYER_MODE=stream" "CGCN_LAYER_MODE=gather" "CGCN_LAYER_MODE=unfused"; do
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-roofline > gpurun_out/r02_bench14_$tag.log 2>&1
  echo "$cfg: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench14_$tag.log | head -1) $(grep -o '"final_loss_sum": [0-9.]*' gpurun_out/r02_bench14_$tag.log)"
done
ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_stream14.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-roofline > gpurun_out/ncu_a.log 2>&1
timeout -k 10 600 python bench.py --workload st --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-roofline 2>&1 | grep -o '"ms_per_step": [0-9.]*' | head -1
