#!/bin/bash
# serving path: T blocking single-query clients -> coalescer -> batch calls, 1M x 768 index
mkdir -p gpurun_out
python tools/make_snapshot.py --out /tmp/snap --raw > gpurun_out/r2q_mk.log 2>&1; tail -1 gpurun_out/r2q_mk.log
gcc -std=c99 -O2 -Iinclude -Iharness harness/coalescer_load.c -Lpgvectorscale_b200 -l:libdiskann_b200.so -Wl,-rpath,$PWD/pgvectorscale_b200 -lpthread -o /tmp/coalescer_load || exit 1
: > gpurun_out/r2q_coalescer.jsonl
for T in 64 256 1024 4096; do
  timeout 300 /tmp/coalescer_load /tmp/snap.raw /tmp/snap_q.f32 $T 32 150 250 10 4096 300 >> gpurun_out/r2q_coalescer.jsonl 2>> gpurun_out/r2q_coalescer.err
done
cat gpurun_out/r2q_coalescer.jsonl; tail -3 gpurun_out/r2q_coalescer.err
