#!/bin/bash
mkdir -p gpurun_out
timeout 2400 compute-sanitizer --tool memcheck --print-limit 10 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_san_full.log 2>&1; grep -E "Invalid|========= +at|passed|failed|ERROR SUMMARY|Address 0x" gpurun_out/r2m_san_full.log | sort | uniq -c | head -30
