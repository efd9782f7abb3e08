#!/bin/bash
# Interleaved timing of several builds of libclothgpu.so on the SAME box (run under gpurun):
#   tools/ab_time.sh LIB [LIB ...]        SHAPES="nx ny K frames;..." selects the shapes
# Builds to compare: tools/build_variant.sh NAME [-DFLAG ...] -> build/variants/NAME.so; an older commit:
#   git archive COMMIT cloth-physics_b200/csrc include | tar -x -C /tmp/old && make -C /tmp/old/cloth-physics_b200/csrc
#   && cp /tmp/old/cloth-physics_b200/csrc/libclothgpu.so build/r1/   (build/ is git-ignored but travels with gpurun)
SHAPES=${SHAPES:-"1024 8192 8 40;16384 16384 1 40;4096 4096 1 100;16384 16384 8 10"}
IFS=';' read -ra SH <<< "$SHAPES"
for rep in 1 2; do
  for shape in "${SH[@]}"; do
    for lib in "$@"; do
      echo -n "$(basename $lib .so | sed 's/libclothgpu/product/') $(basename $(dirname $lib)) : "; CLOTHGPU_LIB=$lib python tools/time_shape.py $shape
    done
  done
done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,clocks_event_reasons.sw_power_cap --format=csv
