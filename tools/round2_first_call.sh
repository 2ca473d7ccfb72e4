#!/bin/bash
# First GPU call of the next round: hardware verdict on everything that was written after this round's GPU minutes
# were spent (bit-exact under CPU emulation only so far), then the A/B timing that decides what becomes the default.
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/round2_first_call.sh'   (about 30 GPU-minutes)
# Output lands in gpurun_out/ (r2_*.log / .json / .ncu-rep).
mkdir -p gpurun_out
# 1. regular parity suite must still be green (default kernels are byte-identical, profiles/r01_sass_signature.txt)
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gpu_tests.log 2>&1; tail -2 gpurun_out/r2_gpu_tests.log
# 2. the opt-in paths: heap engine v2 / controller alternatives (DANN_HEAP_V2=1) and the plain storage layout
DANN_RUN_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_zz_experimental_gpu.py -m gpu -q > gpurun_out/r2_experimental.log 2>&1
tail -3 gpurun_out/r2_experimental.log
# 3. A/B timing on the benchmark fixture at the benchmark's operating point, each alternative alone
python tools/make_snapshot.py --out /tmp/snap --raw > gpurun_out/r2_mk.log 2>&1; tail -1 gpurun_out/r2_mk.log
timeout 600 python tools/hv_ab.py --snap /tmp/snap --L 150 --rescore 250 --steps 20 > gpurun_out/r2_hv_ab.jsonl 2> gpurun_out/r2_hv_ab.err
cat gpurun_out/r2_hv_ab.jsonl
# 4. bench lines: default, and the full HV=1 flavour
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_default.json 2> gpurun_out/r2_bench_default.log
DANN_HEAP_V2=1 timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench_hv1.json 2> gpurun_out/r2_bench_hv1.log
cat gpurun_out/r2_bench_default.json gpurun_out/r2_bench_hv1.json
# 5. one ncu capture of the HV=1 kernel (source-level stalls) for the profile-driven next step
DANN_HEAP_V2=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:dann_search2_kernel -c 1 \
    -o gpurun_out/r2_search2_hv1 python tools/profile_search.py --snap /tmp/snap --L 150 --rescore 250 --steps 2 \
    > gpurun_out/r2_ncu.log 2>&1; tail -2 gpurun_out/r2_ncu.log
# 6. query coalescer: C pthread clients (stand-ins for backends) -> batches; how much of the batch throughput survives
gcc -std=c99 -O2 -Iinclude -Iharness harness/coalescer_load.c -Lpgvectorscale_b200 -l:libdiskann_b200.so \
    -Wl,-rpath,$PWD/pgvectorscale_b200 -lpthread -o /tmp/coalescer_load
for T in 64 256 1024; do
  /tmp/coalescer_load /tmp/snap.raw /tmp/snap_q.f32 $T 64 150 250 10 1024 200 >> gpurun_out/r2_coalescer.jsonl 2>> gpurun_out/r2_coalescer.err
done
cat gpurun_out/r2_coalescer.jsonl
# 7. what the fused expansion's registers cost the rest of the HV=1 kernel: rebuild without it, re-time, restore
DANN_NVCC_DEFINES="-DDANN_HV_NO_FUSED" python -c "from pgvectorscale_b200.build import build_library as b; b(force=True)"
timeout 600 python tools/hv_ab.py --snap /tmp/snap --L 150 --rescore 250 --steps 20 --no-check > gpurun_out/r2_hv_ab_nofused_build.jsonl 2>> gpurun_out/r2_hv_ab.err
python -c "from pgvectorscale_b200.build import build_library as b; b(force=True)"
cat gpurun_out/r2_hv_ab_nofused_build.jsonl
