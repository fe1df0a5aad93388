cd $GRAFT_REPO_ROOT
run() { n=$1; shift; out=$1; shift; timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 296$n bench.py --gpus $n "$@" > gpurun_out/$out 2> gpurun_out/$out.err; grep -o '"ms_per_step": [0-9.]*' gpurun_out/$out | head -2 | tr '\n' ' '; echo " <- $out"; grep -i -E "error|Traceback" gpurun_out/$out.err | head -3; }
B="--no-cpu-baseline --no-gpu-baseline --no-roofline"
for v in layers3 gateoff normnone hic1000000 hic125000; do run 8 r02_bench_wg_8gpu_$v.json --variant $v --steps 10 --warmup 3 $B; done
run 8 r02_bench_st_d512_8gpu.json --workload st --d 512 --steps 3 --warmup 3 $B --no-e2e
run 2 r02_bench_wg_2gpu.json --steps 20 --warmup 5 $B
run 4 r02_bench_st_d512_4gpu.json --workload st --d 512 --steps 3 --warmup 3 $B --no-e2e
run 2 r02_bench_st_d512_2gpu.json --workload st --d 512 --steps 3 --warmup 3 $B --no-e2e
