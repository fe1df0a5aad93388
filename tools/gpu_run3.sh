cd $GRAFT_REPO_ROOT
for cfg in "CGCN_FUSED_LD=0 CGCN_FUSED_GW=16" "CGCN_FUSED_LD=1 CGCN_FUSED_GW=16" "CGCN_FUSED_LD=2 CGCN_FUSED_GW=16" "CGCN_FUSED_LD=1 CGCN_FUSED_GW=8" "CGCN_FUSED_LD=2 CGCN_FUSED_GW=8"; do
  tag=$(echo $cfg | tr ' =' '__')
  env $cfg timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/r02_bench3_$tag.log 2>&1
  echo "$cfg: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench3_$tag.log | head -1) $(grep -o '"final_loss_sum": [0-9.]*' gpurun_out/r02_bench3_$tag.log)"
done
