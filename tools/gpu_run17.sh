cd $GRAFT_REPO_ROOT
run() { n=$1; shift; out=$1; shift; timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n bench.py --gpus $n "$@" > gpurun_out/$out 2> gpurun_out/$out.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/$out | head -2 | tr '\n' ' '; echo " <- $out"; tail -2 gpurun_out/$out.err | cut -c1-200; }
run 8 r02_bench_wg_8gpu.json --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline
run 8 r02_bench_wg_8gpu_rounds2.json --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline --rounds 2 --no-e2e --no-roofline
run 4 r02_bench_wg_4gpu.json --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-baseline
run 8 r02_bench_st_d128_8gpu.json --workload st --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline
run 4 r02_bench_st_d128_4gpu.json --workload st --steps 5 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e
