#!/bin/bash
# Final round-2 capture (one GPU): launch list of a short bench run + --set full of the four per-step kernels of one step.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_r02_final.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-drag-profile --no-mode0 > gpurun_out/launches_r02_final.log 2>&1
ncu --set full --clock-control none --import-source on -k "regex:k_apply_union|k_lbs_union32|k_rotate_sample_shs|k_solve_pipe" -s 8 -c 4 \
    -o /tmp/prof_r02_final -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-drag-profile --no-mode0 > gpurun_out/prof_r02_final.log 2>&1
ncu -i /tmp/prof_r02_final.ncu-rep --page raw --csv > gpurun_out/prof_r02_final_raw.csv
ncu -i /tmp/prof_r02_final.ncu-rep --page source --csv -k regex:k_solve_pipe > gpurun_out/prof_r02_final_solve_source.csv 2>/dev/null
ls -la gpurun_out | tail -5
