cd $GRAFT_REPO_ROOT
timeout -k 10 900 python -m pytest tests/test_gpu_model.py -m gpu -q --timeout 180 -x 2>&1 | tail -40
