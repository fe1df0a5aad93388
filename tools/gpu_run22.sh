cd $GRAFT_REPO_ROOT
timeout -k 10 300 python __graft_entry__.py --smoke 2>&1 | grep -E "smoke|Error|error" | head -5
timeout -k 10 1500 python -m pytest tests/test_gpu_model.py tests/test_gpu_kernels.py tests/test_gpu_rowpartition.py -m gpu -q --timeout 900 2>&1 | tail -5
timeout -k 10 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e 2>&1 | grep -o '"ms_per_step": [0-9.]*\|"avg_launch_us": [0-9.]*' | head -3
