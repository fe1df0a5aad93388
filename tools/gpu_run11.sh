cd $GRAFT_REPO_ROOT
timeout -k 10 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q --timeout 180 -k "fused_layer" 2>&1 | tail -8
python tools/l2_bw.py > gpurun_out/r02_l2_bw.json 2> gpurun_out/l2_err.log; cat gpurun_out/r02_l2_bw.json; tail -3 gpurun_out/l2_err.log
cp gpurun_out/r02_l2_bw.json profiles/r02_l2_bw.json
( time timeout -k 10 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/r02_bench_wg_1gpu.json 2> gpurun_out/bench11_err.log
tail -c 6000 gpurun_out/r02_bench_wg_1gpu.json; tail -5 gpurun_out/bench11_err.log
