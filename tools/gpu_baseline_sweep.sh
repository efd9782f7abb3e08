#!/bin/bash
# One gpurun call: pipe peaks + bench over the workloads and both schedules (no CPU leg).
set -x
mkdir -p gpurun_out
./build/pipe_peaks > gpurun_out/pipe_peaks.json 2>&1
for wl in c1 c1k1 c2 c2k1 c3; do
  python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_${wl}_fused.json 2> gpurun_out/bench_${wl}_fused.err
done
for wl in c1 c1k1; do
  python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu --schedule streaming > gpurun_out/bench_${wl}_streaming.json 2> gpurun_out/bench_${wl}_streaming.err
done
python bench.py --workload c1 --steps 30 --warmup 5 --no-cpu --precision f32 > gpurun_out/bench_c1_f32.json 2>&1
cat gpurun_out/pipe_peaks.json gpurun_out/bench_*.json
