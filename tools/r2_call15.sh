#!/bin/bash
mkdir -p gpurun_out
for i in 1 2; do timeout 900 python -m pytest tests -m gpu -q --deselect tests/test_zz_j_config1_gpu.py > gpurun_out/r2p_gpu_tests_$i.log 2>&1; echo "full suite run $i: $(tail -1 gpurun_out/r2p_gpu_tests_$i.log)"; done
timeout 600 compute-sanitizer --tool synccheck --print-limit 6 python -m pytest tests/test_zz_f_fuzz_gpu.py -m gpu -x -q -k "batch_and_scan_equal_oracle and (1 or 2 or 3)" > gpurun_out/r2p_synccheck.log 2>&1; grep -E "Barrier|Divergent|========= +at|passed|failed|ERROR SUMMARY" gpurun_out/r2p_synccheck.log | sort | uniq -c | sort -rn | head -8
timeout 600 python -m pytest tests/test_zz_j_config1_gpu.py -m gpu -q > gpurun_out/r2p_config1.log 2>&1; tail -2 gpurun_out/r2p_config1.log
python tools/make_snapshot.py --out /tmp/snap > gpurun_out/r2p_mk.log 2>&1
timeout 600 python tools/lean_ab.py --snap /tmp/snap --L 150 --rescore 250 --steps 6 --batches 1024,4096 > gpurun_out/r2p_lean_ab.jsonl 2> gpurun_out/r2p_lean_ab.err; cut -c1-200 gpurun_out/r2p_lean_ab.jsonl
timeout 900 python tools/profile_big.py --n 8000000 --L 800 --rescore 800 --batch 4096 --steps 4 --check 64 > gpurun_out/r2p_8m.json 2> gpurun_out/r2p_8m.log; cut -c1-900 gpurun_out/r2p_8m.json
