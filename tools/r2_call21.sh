#!/bin/bash
# final single-GPU evidence: full suite, bench at the 50M headline (both arms), coalescer throughput
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2v_gpu_tests.log 2>&1; echo "full suite: $(tail -1 gpurun_out/r2v_gpu_tests.log)"
timeout 1500 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2v_bench_50m_ref.json 2> gpurun_out/r2v_bench_50m_ref.log; grep -E "^\[bench|rror" gpurun_out/r2v_bench_50m_ref.log | tail -4; cut -c1-400 gpurun_out/r2v_bench_50m_ref.json
timeout 1500 python bench.py > gpurun_out/r2v_bench_50m.json 2> gpurun_out/r2v_bench_50m.log; grep -E "^\[bench|rror" gpurun_out/r2v_bench_50m.log | tail -12; cut -c1-600 gpurun_out/r2v_bench_50m.json
python tools/make_snapshot.py --out /tmp/snap --raw > gpurun_out/r2v_mk.log 2>&1
gcc -std=c99 -O2 -Iinclude -Iharness harness/coalescer_load.c -Lpgvectorscale_b200 -l:libdiskann_b200.so -Wl,-rpath,$PWD/pgvectorscale_b200 -lpthread -o /tmp/coalescer_load && {
  : > gpurun_out/r2v_coalescer.jsonl
  for T in 64 256 1024 4096; do timeout 200 /tmp/coalescer_load /tmp/snap.raw /tmp/snap_q.f32 $T 24 150 250 10 4096 300 >> gpurun_out/r2v_coalescer.jsonl 2>> gpurun_out/r2v_coalescer.err; done
  cat gpurun_out/r2v_coalescer.jsonl; tail -2 gpurun_out/r2v_coalescer.err; }
