#!/bin/bash
# Interleaved timing of the live-stick export of several builds of the library on the SAME box (run under gpurun):
#   tools/ab_export.sh LIB [LIB ...]     SHAPES="nx ny reps;..." selects the cloths
SHAPES=${SHAPES:-"4096 4096 200;171 36 500"}
IFS=';' read -ra SH <<< "$SHAPES"
for rep in 1 2; do
  for shape in "${SH[@]}"; do
    for lib in "$@"; do
      echo -n "$(basename $lib .so) : "; CLOTHGPU_LIB=$lib python tools/time_export.py $shape
    done
  done
done
