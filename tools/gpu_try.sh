#!/bin/bash
# tests, then bench variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/pytest_gpu.log
cat gpurun_out/pytest_gpu.log
for v in 0 1 2; do
  for wl in c1 c1k1; do
    CLOTHGPU_FUSED_VARIANT=$v timeout 300 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_${wl}_v$v.json 2> gpurun_out/bench_${wl}_v$v.err
    python - gpurun_out/bench_${wl}_v$v.json <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "ms/step", round(d["ms_per_step"],4), "Grelax/s", round(d["value"]/1e9,1), "hbm frac", round(d["roofline"]["frac"],4))
except Exception as e:
    print("ERR", sys.argv[1], e)
PY
  done
done
