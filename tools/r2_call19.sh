#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; timeout 600 python -m pytest "$@" -m gpu -x -q > gpurun_out/r2t_$name.log 2>&1; echo "$name: $(tail -1 gpurun_out/r2t_$name.log)"; }
run A tests/test_zz_c_harness_gpu.py tests/test_zz_d_build_small_gpu.py tests/test_zz_e_coalescer_gpu.py tests/test_zz_f_fuzz_gpu.py
run B tests/test_gpu_build.py tests/test_gpu_parity.py tests/test_gpu_properties_1m.py tests/test_zz_f_fuzz_gpu.py
run C tests/test_zz_d_build_small_gpu.py tests/test_zz_f_fuzz_gpu.py
run D tests/test_zz_e_coalescer_gpu.py tests/test_zz_f_fuzz_gpu.py
run E tests/test_gpu_properties_1m.py tests/test_zz_d_build_small_gpu.py tests/test_zz_f_fuzz_gpu.py
nvidia-smi --query-gpu=name,driver_version,ecc.errors.uncorrected.volatile.total,ecc.errors.corrected.volatile.total,retired_pages.pending --format=csv
