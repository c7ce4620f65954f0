#!/bin/bash
# Round-2 profiling recipe (run under gpurun, one GPU).  Outputs go to gpurun_out/ (<= 64 MiB come back: the .ncu-rep files are
# exported to CSV on the box and removed); summaries are copied to profiles/.
set -x
mkdir -p gpurun_out
TAG=${1:-r02}
# (1) every launch of a short bench run with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-drag-profile > gpurun_out/launches_$TAG.log 2>&1
# (2) full sections of the per-step kernels of the reported configuration (lbs_mode 3): one step after set-up and warm-up
ncu --set full --clock-control none --import-source on -k "regex:k_apply_union|k_lbs_union32|k_rotate_sample_shs|k_solve_smem" -s 8 -c 4 \
    -o /tmp/prof_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-drag-profile > gpurun_out/prof_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv -k regex:k_apply_union > gpurun_out/prof_${TAG}_apply_source.csv 2>/dev/null
# (3) full sections of the set-up / stroke-end kernels (stages (a) and (b))
ncu --set full --clock-control none -k "regex:k_knn_tile|k_grid_eval|k_footprint|k_segsort|k_gunion_build|k_sunion_build|k_cell_hist|k_permute|k_fps_pruned" -c 16 \
    -o /tmp/prof_setup_$TAG -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-drag-profile > gpurun_out/prof_setup_$TAG.log 2>&1
ncu -i /tmp/prof_setup_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_setup_${TAG}_raw.csv
ls -la gpurun_out
