cd $GRAFT_REPO_ROOT
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 180 -k "fused_layer" 2>&1 | tail -3
( time timeout -k 10 1200 python bench.py --workload st --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/r02_bench_st_d128_1gpu.json 2> gpurun_out/bench12_err.log
tail -c 3000 gpurun_out/r02_bench_st_d128_1gpu.json; tail -5 gpurun_out/bench12_err.log
( time timeout -k 10 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/r02_bench_wg_2gpu.json 2> gpurun_out/bench12b_err.log
tail -c 2500 gpurun_out/r02_bench_wg_2gpu.json; tail -5 gpurun_out/bench12b_err.log
