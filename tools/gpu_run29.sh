#!/bin/bash
# multi-GPU parity tests on a 2-GPU lease (sharded pass, native NCCL entry points, row-partitioned graph with peer loads)
mkdir -p gpurun_out
timeout -k 5 170 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_rowpartition.py -m gpu -q -rs > gpurun_out/r02_pytest_2gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_2gpu_final.log
