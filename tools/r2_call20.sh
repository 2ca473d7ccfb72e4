#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m pytest "$@" -m gpu -x -q > gpurun_out/r2u_$name.log 2>&1; echo "$name: $(tail -1 gpurun_out/r2u_$name.log)"; }
run A tests/test_zz_c_harness_gpu.py tests/test_zz_d_build_small_gpu.py tests/test_zz_e_coalescer_gpu.py tests/test_zz_f_fuzz_gpu.py
run C tests/test_zz_d_build_small_gpu.py tests/test_zz_f_fuzz_gpu.py
for i in 1 2; do timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2u_full_$i.log 2>&1; echo "full suite run $i: $(tail -1 gpurun_out/r2u_full_$i.log)"; done
