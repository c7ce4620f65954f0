#!/bin/bash
# Round-1 profiling recipe (run under gpurun, one GPU).  Outputs go to gpurun_out/; summaries are copied to profiles/.
set -x
mkdir -p gpurun_out
W=${1:-shells6m}
TAG=${2:-r01d}
# (1) every launch of a short bench run with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --workload $W --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/launches_$TAG.log 2>&1
# (2) full sections for the per-step kernels (one step's worth, after set-up and warm-up)
ncu --set full --clock-control none --import-source on -k "regex:k_lbs_tiles|k_lbs_points|k_fit_gaussians|k_rotate_sample_shs|k_solve" -s 18 -c 6 \
    -o gpurun_out/prof_$TAG -f python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_$TAG.log 2>&1
ls -la gpurun_out
