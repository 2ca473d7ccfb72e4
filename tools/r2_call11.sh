#!/bin/bash
mkdir -p gpurun_out
for f in test_gpu_build test_gpu_parity test_gpu_properties_1m; do
  timeout 600 python -m pytest tests/$f.py tests/test_zz_f_fuzz_gpu.py -m gpu -x -q > gpurun_out/r2l_$f.log 2>&1; echo "$f: $(tail -1 gpurun_out/r2l_$f.log)"
done
for f in test_gpu_build test_gpu_parity; do
  timeout 900 compute-sanitizer --tool memcheck --print-limit 6 python -m pytest tests/$f.py tests/test_zz_f_fuzz_gpu.py -m gpu -x -q -k "not properties" > gpurun_out/r2l_san_$f.log 2>&1; echo "san $f:"; grep -E "Invalid|========= +at|passed|failed|ERROR SUMMARY" gpurun_out/r2l_san_$f.log | sort | uniq -c | head -12
done
