#!/bin/bash
# bench a list of fused-kernel variants on a list of workloads: VARIANTS="0 3 4" WLS="c1 c2k1" tools/gpu_variants.sh
mkdir -p gpurun_out
for v in ${VARIANTS:-0 3 4}; do
  for wl in ${WLS:-c1 c1k1}; do
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
