#!/bin/bash
mkdir -p gpurun_out
for c in c5f7110 bb93a1a; do
  (cd _bisect/$c && timeout 600 python -m pytest tests -m gpu -x -q > ../../gpurun_out/r2s_bisect_$c.log 2>&1); echo "$c: $(tail -1 gpurun_out/r2s_bisect_$c.log)"; grep -E "^E  .*error" gpurun_out/r2s_bisect_$c.log | head -2 | cut -c1-300
done
