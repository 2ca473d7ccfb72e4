#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_gpu_tests.log 2>&1; tail -3 gpurun_out/r2h_gpu_tests.log
timeout 900 python tools/profile_big.py --n 8000000 --L 800 --rescore 800 --batch 4096 --steps 6 --check 64 > gpurun_out/r2h_8m.json 2> gpurun_out/r2h_8m.log; tail -2 gpurun_out/r2h_8m.log; cat gpurun_out/r2h_8m.json
timeout 1500 ncu --set full --import-source on --clock-control none -k regex:dann_search3_kernel -s 1 -c 1 \
    -o /tmp/prof_8m python tools/profile_big.py --n 8000000 --L 800 --rescore 800 --batch 4096 --steps 1 > gpurun_out/r2h_ncu.log 2>&1; tail -2 gpurun_out/r2h_ncu.log
python tools/ncu_summary.py /tmp/prof_8m.ncu-rep 90 > gpurun_out/r2h_search3_8m_summary.txt 2>&1
ncu -i /tmp/prof_8m.ncu-rep --page raw --csv > gpurun_out/r2h_search3_8m_raw.csv 2>/dev/null
head -32 gpurun_out/r2h_search3_8m_summary.txt
timeout 900 python bench.py --mode scan --n 1000000 > gpurun_out/r2h_scan_latency.json 2> gpurun_out/r2h_scan_latency.log; tail -2 gpurun_out/r2h_scan_latency.log; cat gpurun_out/r2h_scan_latency.json
