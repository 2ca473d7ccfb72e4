#!/bin/bash
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --print-limit 8 python -m pytest tests/test_zz_f_fuzz_gpu.py tests/test_zz_plain_gpu.py -m gpu -x -q > gpurun_out/r2i_sanitizer.log 2>&1; grep -E "Invalid|========= +at|passed|failed|ERROR SUMMARY" gpurun_out/r2i_sanitizer.log | sort | uniq -c | head -20
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2i_gpu_tests.log 2>&1; tail -6 gpurun_out/r2i_gpu_tests.log
DANN_SEARCH_KERNEL=3 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_zz_f_fuzz_gpu.py -m gpu -q > gpurun_out/r2i_gpu_tests_lean_forced.log 2>&1; tail -3 gpurun_out/r2i_gpu_tests_lean_forced.log
