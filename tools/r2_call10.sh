#!/bin/bash
mkdir -p gpurun_out
CUDA_LAUNCH_BLOCKING=1 timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_blocking.log 2>&1; grep -E "^E  |passed|failed" gpurun_out/r2k_blocking.log | head -8
timeout 1500 compute-sanitizer --tool memcheck --print-limit 6 python -m pytest tests/test_zz_c_harness_gpu.py tests/test_zz_d_build_small_gpu.py tests/test_zz_e_coalescer_gpu.py tests/test_zz_f_fuzz_gpu.py -m gpu -x -q > gpurun_out/r2k_sanitizer_seq.log 2>&1; grep -E "Invalid|========= +at|passed|failed|ERROR SUMMARY" gpurun_out/r2k_sanitizer_seq.log | sort | uniq -c | head -20
