#!/bin/bash
# Round-2 GPU call: full parity suite on hardware, then bench.py at configs[1] and at the 50M headline.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_gpu_tests.log 2>&1; tail -3 gpurun_out/r2e_gpu_tests.log
timeout 900 python bench.py --n 1000000 --steps 10 --warmup 3 > gpurun_out/r2e_bench_1m.json 2> gpurun_out/r2e_bench_1m.log; tail -4 gpurun_out/r2e_bench_1m.log; cat gpurun_out/r2e_bench_1m.json | cut -c1-1500
timeout 1500 python bench.py --steps 10 --warmup 3 > gpurun_out/r2e_bench_50m.json 2> gpurun_out/r2e_bench_50m.log; tail -25 gpurun_out/r2e_bench_50m.log; cat gpurun_out/r2e_bench_50m.json | cut -c1-3000
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
