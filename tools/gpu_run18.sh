cd $GRAFT_REPO_ROOT
B="--no-cpu-baseline --no-gpu-baseline"
for v in layers3 gateoff normnone hic1000000 hic125000; do
  timeout -k 10 600 python bench.py --variant $v --steps 10 --warmup 3 $B --no-roofline > gpurun_out/r02_bench_wg_1gpu_$v.json 2> gpurun_out/v_$v.err
  echo "$v: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_wg_1gpu_$v.json | head -2 | tr '\n' ' ') $(grep -o '"value": [0-9.]*' gpurun_out/r02_bench_wg_1gpu_$v.json | head -1)"; tail -1 gpurun_out/v_$v.err | cut -c1-160
done
timeout -k 10 600 python bench.py --workload c1 --steps 200 --warmup 10 $B > gpurun_out/r02_bench_c1_1gpu.json 2> gpurun_out/c1.err
echo "c1: $(grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_c1_1gpu.json | head -2 | tr '\n' ' ')"; tail -1 gpurun_out/c1.err | cut -c1-160
ncu --query-metrics 2>/dev/null | grep -i -E "tensor|tmem|utc" | head -40 > gpurun_out/r02_ncu_tensor_metrics_available.txt
N="--no-cpu-baseline --no-gpu-baseline --no-e2e --no-roofline"
ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_wg.csv python bench.py --steps 1 --warmup 3 $N > gpurun_out/ncu_a.log 2>&1
ncu --nvtx --nvtx-include "timed/" -k regex:"fused_layer|gemm_rowpanel_tc|gemm_gram_tc|spmm_pattern" --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__cycles_elapsed.max --clock-control none --csv --log-file gpurun_out/r02_kernel_metrics_wg.csv python bench.py --steps 1 --warmup 3 $N > gpurun_out/ncu_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_layer -s 44 -c 3 -f -o gpurun_out/r02_fused_final python bench.py --steps 1 --warmup 3 $N > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out | tail -12
