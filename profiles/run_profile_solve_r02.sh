#!/bin/bash
# ncu source-level capture of the solver kernel (one launch of a steady drag step), exported to CSV on the box.
set -x
mkdir -p gpurun_out
TAG=${1:-r02_pipe}
KERN=${2:-k_solve_pipe}
ncu --set full --clock-control none --import-source on -k "regex:$KERN" -s 5 -c 1 \
    -o /tmp/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-drag-profile > gpurun_out/prof_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
ls -la gpurun_out | tail -5
