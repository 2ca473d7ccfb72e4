#!/bin/bash
# last minutes of the round's GPU budget: bench.py's glue after the final edits (per-n sweep, -march=native CPU arm,
# operator key) on configs[1], both arms
mkdir -p gpurun_out
timeout 80 python bench.py --n 1000000 --steps 5 --warmup 3 > gpurun_out/r2z_bench_1m.json 2> gpurun_out/r2z_bench_1m.log; tail -3 gpurun_out/r2z_bench_1m.log; cut -c1-300 gpurun_out/r2z_bench_1m.json; echo
timeout 80 python bench.py --impl reference --n 1000000 --steps 3 --warmup 1 > gpurun_out/r2z_ref_1m.json 2> gpurun_out/r2z_ref_1m.log; tail -2 gpurun_out/r2z_ref_1m.log; cut -c1-300 gpurun_out/r2z_ref_1m.json
