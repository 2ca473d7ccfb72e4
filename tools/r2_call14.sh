#!/bin/bash
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool initcheck --print-limit 12 python -m pytest tests/test_zz_f_fuzz_gpu.py -m gpu -x -q -k "batch_and_scan_equal_oracle and (3 or 6)" > gpurun_out/r2o_initcheck.log 2>&1; grep -E "Uninitialized|========= +at|passed|failed|ERROR SUMMARY" gpurun_out/r2o_initcheck.log | sort | uniq -c | sort -rn | head -30
timeout 600 compute-sanitizer --tool synccheck --print-limit 12 python -m pytest tests/test_zz_f_fuzz_gpu.py -m gpu -x -q -k "batch_and_scan_equal_oracle and (3 or 6)" > gpurun_out/r2o_synccheck.log 2>&1; grep -E "Barrier|Divergent|========= +at|passed|failed|ERROR SUMMARY" gpurun_out/r2o_synccheck.log | sort | uniq -c | sort -rn | head -20
