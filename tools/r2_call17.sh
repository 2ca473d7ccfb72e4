#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q --deselect tests/test_zz_j_config1_gpu.py > gpurun_out/r2r_gpu_tests.log 2>&1; grep -E "^E  .*error|passed|failed" gpurun_out/r2r_gpu_tests.log | head -5
