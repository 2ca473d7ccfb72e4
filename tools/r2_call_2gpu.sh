#!/bin/bash
# two GPUs: the in-process group (dann_group_*) against the oracle, then bench.py under torchrun (NCCL gather + check)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 600 python -m pytest tests/test_zz_i_group_gpu.py -m gpu -q > gpurun_out/r2j_group_tests.log 2>&1; tail -3 gpurun_out/r2j_group_tests.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --n 1000000 --batch 4096 --steps 10 --warmup 3 > gpurun_out/r2j_bench_2gpu_1m.json 2> gpurun_out/r2j_bench_2gpu_1m.log; grep -E "^\[bench|rror|NCCL" gpurun_out/r2j_bench_2gpu_1m.log | tail -12; cut -c1-1200 gpurun_out/r2j_bench_2gpu_1m.json
