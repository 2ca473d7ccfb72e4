#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_zz_f_fuzz_gpu.py -m gpu -x -q -k "batch_and_scan_equal_oracle and (3 or 2)" > gpurun_out/r2i_sanitizer.log 2>&1; grep -E "Invalid|at 0x|by 0x|Address|========= +at|passed|failed|ERROR SUMMARY" gpurun_out/r2i_sanitizer.log | head -40
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2i_gpu_tests.log 2>&1; tail -8 gpurun_out/r2i_gpu_tests.log
