set -x
cd $GRAFT_REPO_ROOT
ncu --nvtx --nvtx-include "timed/" --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_fused1.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_layer -s 46 -c 4 -f -o gpurun_out/r02_fused1 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out
