#!/bin/bash
# bench.py at the 50M headline (ours, then the reference arm), with wall-clock of each.
mkdir -p gpurun_out
timeout 1800 python bench.py --steps 10 --warmup 3 > gpurun_out/r2f_bench_50m.json 2> gpurun_out/r2f_bench_50m.log; grep -E "^\[bench|rror" gpurun_out/r2f_bench_50m.log | tail -30; cat gpurun_out/r2f_bench_50m.json | cut -c1-2500
timeout 1800 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2f_bench_50m_ref.json 2> gpurun_out/r2f_bench_50m_ref.log; grep -E "^\[bench|Elapsed|Maximum resident|Error|error" gpurun_out/r2f_bench_50m_ref.log | tail -20; cat gpurun_out/r2f_bench_50m_ref.json | cut -c1-2000
