cd $GRAFT_REPO_ROOT
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29688 bench.py --gpus 8 --workload st --d-model 512 --steps 3 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-roofline --no-e2e > gpurun_out/r02_bench_st_d512_8gpu.json 2> gpurun_out/st512.err
grep -o '"ms_per_step": [0-9.]*' gpurun_out/r02_bench_st_d512_8gpu.json | head -1; grep -i -E "error|Traceback" gpurun_out/st512.err | head -3
