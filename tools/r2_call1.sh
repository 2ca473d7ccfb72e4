#!/bin/bash
# Round-2 first GPU call: hardware verdict on the paths that had only run under emulation, HV flag A/B,
# free-memory probe, and a B=4096 baseline of the round-1 kernels on a mid-size index.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,memory.used,memory.free --format=csv > gpurun_out/r2a_mem.txt 2>&1
python - >> gpurun_out/r2a_mem.txt 2>&1 <<'PY'
import torch, os
f, t = torch.cuda.mem_get_info()
print("mem_get_info free/total bytes", f, t)
print("cpus", os.cpu_count(), "affinity", len(os.sched_getaffinity(0)))
try: print("cpu.max", open("/sys/fs/cgroup/cpu.max").read().strip())
except Exception as e: print("cpu.max n/a", e)
import shutil; print("disk /tmp", shutil.disk_usage("/tmp")); print("memtotal", open("/proc/meminfo").readline().strip())
PY
cat gpurun_out/r2a_mem.txt
DANN_RUN_EXPERIMENTAL=1 timeout 900 python -m pytest tests/test_zz_experimental_gpu.py -m gpu -q > gpurun_out/r2a_experimental.log 2>&1
tail -5 gpurun_out/r2a_experimental.log
python tools/make_snapshot.py --out /tmp/snap > gpurun_out/r2a_mk.log 2>&1; tail -1 gpurun_out/r2a_mk.log
timeout 600 python tools/hv_ab.py --snap /tmp/snap --L 150 --rescore 250 --steps 20 > gpurun_out/r2a_hv_ab.jsonl 2> gpurun_out/r2a_hv_ab.err
cat gpurun_out/r2a_hv_ab.jsonl
# B=4096 on 4M nodes: the round-1 two-warp kernel (7 queries/SM) and the single-warp one (12/SM)
timeout 900 python tools/large_recall.py --n 4000000 --batch 4096 --check 32 --cpu-sample 32 > gpurun_out/r2a_4m_pairs.json 2> gpurun_out/r2a_4m_pairs.err
tail -3 gpurun_out/r2a_4m_pairs.err
DANN_SEARCH_KERNEL=1 timeout 900 python tools/large_recall.py --n 4000000 --batch 4096 --check 32 --cpu-sample 32 > gpurun_out/r2a_4m_single.json 2> gpurun_out/r2a_4m_single.err
tail -3 gpurun_out/r2a_4m_single.err
