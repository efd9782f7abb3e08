#!/bin/bash
# plain run first, then one ncu --set full capture of the fused kernel (1 GPU)
mkdir -p gpurun_out
WL=${WL:-c1}
NAME=${NAME:-prof}
python bench.py --workload $WL --steps 6 --warmup 3 --no-cpu ${BENCH_ARGS:-} > gpurun_out/plain_$NAME.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:${KREGEX:-fused_frame} -s 4 -c 1 -f -o gpurun_out/$NAME \
    python bench.py --workload $WL --steps 6 --warmup 3 --no-cpu ${BENCH_ARGS:-} > gpurun_out/ncu_$NAME.log 2>&1
tail -3 gpurun_out/ncu_$NAME.log
