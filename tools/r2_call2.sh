#!/bin/bash
# Round-2 second GPU call: the lean warp-per-query kernel on hardware - parity suite, A/B against the round-1 kernel,
# one ncu capture at batch 4096.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_gpu_tests.log 2>&1; tail -3 gpurun_out/r2b_gpu_tests.log
python tools/make_snapshot.py --out /tmp/snap > gpurun_out/r2b_mk.log 2>&1; tail -1 gpurun_out/r2b_mk.log
timeout 900 python tools/lean_ab.py --snap /tmp/snap --L 150 --rescore 250 --steps 8 > gpurun_out/r2b_lean_ab.jsonl 2> gpurun_out/r2b_lean_ab.err
cat gpurun_out/r2b_lean_ab.jsonl; tail -3 gpurun_out/r2b_lean_ab.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dann_search3_kernel -c 1 \
    -o gpurun_out/r2b_search3_b4096 python tools/profile_search.py --snap /tmp/snap --L 150 --rescore 250 --steps 2 --batch 4096 \
    > gpurun_out/r2b_ncu.log 2>&1; tail -2 gpurun_out/r2b_ncu.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dann_search3_kernel -c 1 \
    -o gpurun_out/r2b_search3_b1024 python tools/profile_search.py --snap /tmp/snap --L 150 --rescore 250 --steps 2 --batch 1024 \
    > gpurun_out/r2b_ncu2.log 2>&1; tail -2 gpurun_out/r2b_ncu2.log
