cd $GRAFT_REPO_ROOT
timeout -k 10 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_sharded.py -m gpu -q --timeout 400 -k "module_api_training or three_epoch or chromosome_sharded" 2>&1 | tail -15
( time timeout -k 10 1200 python bench.py --workload st --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/r02_bench_st_d128_1gpu.json 2> gpurun_out/bench13_err.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_st_d128_1gpu.json | head -2; tail -3 gpurun_out/bench13_err.log
( time timeout -k 10 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload st --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/r02_bench_st_d128_2gpu.json 2> gpurun_out/bench13b_err.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_st_d128_2gpu.json | head -2; tail -3 gpurun_out/bench13b_err.log
( time timeout -k 10 1200 python bench.py --workload st --d 512 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline ) > gpurun_out/r02_bench_st_d512_1gpu.json 2> gpurun_out/bench13c_err.log
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_st_d512_1gpu.json | head -2; tail -3 gpurun_out/bench13c_err.log
