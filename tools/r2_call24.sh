#!/bin/bash
# final evidence of round 2: the 50M headline line (builder fixed: device-input calls ordered after the caller's
# default-stream work), then the whole -m gpu suite
mkdir -p gpurun_out
timeout 420 python bench.py > gpurun_out/r2y_bench_50m.json 2> gpurun_out/r2y_bench_50m.log; grep -E "^\[bench|rror|vamana" gpurun_out/r2y_bench_50m.log | tail -22; cut -c1-500 gpurun_out/r2y_bench_50m.json; echo
timeout 260 python -m pytest tests -m gpu -q -x > gpurun_out/r2y_gpu_tests.log 2>&1; echo "full suite: $(tail -1 gpurun_out/r2y_gpu_tests.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/r2y_gpu_tests.log | head -5
