#!/bin/bash
# usage: tools/gpu_ncu_counts.sh VARIANT WORKLOAD KREGEX -> instruction counts / time of one launch of the variant's kernel
cp cloth-physics_b200/csrc/libclothgpu.so /tmp/libclothgpu.orig2.so
cp build/variants/$1.so cloth-physics_b200/csrc/libclothgpu.so
ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_fp64.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:$3 -s 4 -c 1 --csv python bench.py --workload $2 --steps 6 --warmup 3 --no-cpu 2>/dev/null | grep -E "^\"[0-9]" | awk -F'","' '{print $(NF-2), $(NF)}' | tr -d '"' | sed "s/^/$1 $2 /"
cp /tmp/libclothgpu.orig2.so cloth-physics_b200/csrc/libclothgpu.so
