#!/bin/bash
# round 2, after the bucketed inserted-set + early code-row request in the lean kernel: full suite, the 8M hash-flavour
# profile at the 50M operating point (round-2 reference: 43.9 ms per 4096-query launch), the 50M headline line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2w_gpu_tests.log 2>&1; echo "full suite: $(tail -1 gpurun_out/r2w_gpu_tests.log)"
grep -E "^(FAILED|ERROR)" gpurun_out/r2w_gpu_tests.log | head -10
timeout 400 python tools/profile_big.py --n 8000000 --L 800 --rescore 800 --batch 4096 --steps 4 --check 64 > gpurun_out/r2w_profile_8m.json 2> gpurun_out/r2w_profile_8m.log; cut -c1-900 gpurun_out/r2w_profile_8m.json; echo
timeout 1200 python bench.py > gpurun_out/r2w_bench_50m.json 2> gpurun_out/r2w_bench_50m.log; grep -E "^\[bench|rror" gpurun_out/r2w_bench_50m.log | tail -16; cut -c1-700 gpurun_out/r2w_bench_50m.json
