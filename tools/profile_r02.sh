#!/bin/bash
# Round-2 profile captures; run on the GPU box through gpurun from the repo root.  Each ncu pass follows a plain run of
# the same command that exited 0.  Outputs land in gpurun_out/ (summaries are made here by tools/profile_summaries.py).
set -x
python bench.py --steps 5 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_prof_plain.json 2> gpurun_out/r02_prof_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:hsl_gait_pipe_kernel -s 4 -c 1 -f -o gpurun_out/r02_pipe_full \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_ncu_pipe.log 2>&1
python tools/profile_entries.py > gpurun_out/r02_entries_plain.txt 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:'hsl_frames_kernel|hsl_forces_kernel|hsl_fall_kernel|hsl_topk_coop_kernel|hsl_transpose_kernel' \
    -c 12 -f -o gpurun_out/r02_entries_full python tools/profile_entries.py > gpurun_out/r02_ncu_entries.log 2>&1
ls -la gpurun_out/*.ncu-rep
