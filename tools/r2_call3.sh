#!/bin/bash
# Round-2 third GPU call: lean kernel with staged pushes + L2 prefetch - A/B, ncu summary (text only: the .ncu-rep
# with sources is too large to bring back), and the deep-heap case (4M nodes, batch 4096, L=300).
mkdir -p gpurun_out
python tools/make_snapshot.py --out /tmp/snap > gpurun_out/${TAG:-r2c}_mk.log 2>&1; tail -1 gpurun_out/${TAG:-r2c}_mk.log
timeout 900 python tools/lean_ab.py --snap /tmp/snap --L 150 --rescore 250 --steps 8 --batches 1024,4096,8192 > gpurun_out/${TAG:-r2c}_lean_ab.jsonl 2> gpurun_out/${TAG:-r2c}_lean_ab.err
cat gpurun_out/${TAG:-r2c}_lean_ab.jsonl; tail -3 gpurun_out/${TAG:-r2c}_lean_ab.err
timeout 600 ncu --set full --import-source on --clock-control none -k regex:dann_search3_kernel -c 1 \
    -o /tmp/prof_search3_b4096 python tools/profile_search.py --snap /tmp/snap --L 150 --rescore 250 --steps 2 --batch 4096 \
    > gpurun_out/${TAG:-r2c}_ncu.log 2>&1; tail -2 gpurun_out/${TAG:-r2c}_ncu.log
python tools/ncu_summary.py /tmp/prof_search3_b4096.ncu-rep 70 > gpurun_out/${TAG:-r2c}_search3_b4096_summary.txt 2>&1
ncu -i /tmp/prof_search3_b4096.ncu-rep --page raw --csv > gpurun_out/${TAG:-r2c}_search3_b4096_raw.csv 2>/dev/null
ls -la /tmp/prof_search3_b4096.ncu-rep
head -40 gpurun_out/${TAG:-r2c}_search3_b4096_summary.txt
timeout 900 python tools/large_recall.py --n 4000000 --batch 4096 --check 32 --cpu-sample 32 > gpurun_out/${TAG:-r2c}_4m_lean.json 2> gpurun_out/${TAG:-r2c}_4m_lean.err
grep large_recall gpurun_out/${TAG:-r2c}_4m_lean.err | tail -6
