#!/bin/bash
# usage (on the GPU box): tools/gpu_sweep_variants.sh "v1 v2 ..." "wl1 wl2 ..." [pytest -k expression]
# For each build/variants/V.so: install it as the product library, run the wave parity tests, then bench the workloads.
mkdir -p gpurun_out
cp cloth-physics_b200/csrc/libclothgpu.so /tmp/libclothgpu.orig.so
for v in $1; do
  cp build/variants/$v.so cloth-physics_b200/csrc/libclothgpu.so
  if [ -n "$3" ]; then
    timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$3" 2>&1 | tail -2 | tr '\n' ' '; echo " [$v parity]"
  fi
  for wl in $2; do
    timeout 300 python bench.py --workload $wl --steps 60 --warmup 10 --no-cpu ${BENCH_ARGS:-} > gpurun_out/var_${v}_${wl}.json 2> gpurun_out/var_${v}_${wl}.err
    python - gpurun_out/var_${v}_${wl}.json $v $wl <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[2], sys.argv[3], "ms/step", round(d["ms_per_step"],4), "Grelax/s", round(d["value"]/1e9,1), d["roofline"].get("kernel"))
except Exception as e:
    print("ERR", sys.argv[1], e)
PY
  done
done
cp /tmp/libclothgpu.orig.so cloth-physics_b200/csrc/libclothgpu.so
