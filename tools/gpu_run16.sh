cd $GRAFT_REPO_ROOT
CGCN_FUSED_EPI=4 CGCN_FUSED_HEAD=1 ncu --set full --clock-control none --import-source on -k regex:"fused_layer_kernel<2, 4" -s 10 -c 1 -f -o gpurun_out/r02_head16 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-roofline > gpurun_out/ncu_b.log 2>&1
tail -3 gpurun_out/ncu_b.log
