#!/bin/bash
# N-GPU checks: NCCL strip parity test, then strip and ensemble benches.  usage: N=2 WLS="c4s c1" tools/gpu_multi.sh
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "multi_gpu" 2>&1 | tail -5
for wl in ${WLS:-c4s}; do
  for n in ${NS:-1 $N}; do
    if [ $n = 1 ]; then
      timeout 600 python bench.py --workload $wl --steps ${STEPS:-20} --warmup 5 --no-cpu > gpurun_out/bench_${wl}_n$n.json 2> gpurun_out/bench_${wl}_n$n.err
    else
      timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --workload $wl --steps ${STEPS:-20} --warmup 5 --no-cpu > gpurun_out/bench_${wl}_n$n.json 2> gpurun_out/bench_${wl}_n$n.err
    fi
    python - gpurun_out/bench_${wl}_n$n.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith("{")][-1])
    print(sys.argv[1], "n_gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"],4), "Grelax/s", round(d["value"]/1e9,1), "hbm frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["value"]/1e9,1))
except Exception as e:
    print("ERR", sys.argv[1], e); print(open(sys.argv[1].replace(".json",".err")).read()[-1500:])
PY
  done
done
