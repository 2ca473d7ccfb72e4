#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench_50m.json 2> gpurun_out/r2g_bench_50m.log; grep -E "^\[bench|rror" gpurun_out/r2g_bench_50m.log | tail -30; cat gpurun_out/r2g_bench_50m.json | cut -c1-3500
