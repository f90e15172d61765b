#!/bin/bash
# --set full captures of the shipped walk kernel (kept graph with triangle Blooms and hub-pair filter, scale 23) and of the two
# triple-window kernels (0.5 M walks), each after a plain run of the same command.  Run under gpurun from the repository root.
set -u
python tools/profile_target.py --scale 23 --prepared --no-extras --reps 1 > gpurun_out/r2_final_prof_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:node2vec_walk -c 2 -o gpurun_out/r2_final_walk -f \
    python tools/profile_target.py --scale 23 --prepared --no-extras --reps 1 > gpurun_out/r2_final_prof_ncu.log 2>&1
python tools/windows_profile_target.py > gpurun_out/r2_final_win_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:windows_kernel -c 4 -o gpurun_out/r2_final_win -f \
    python tools/windows_profile_target.py > gpurun_out/r2_final_win_ncu.log 2>&1
ls -la gpurun_out/r2_final_walk.ncu-rep gpurun_out/r2_final_win.ncu-rep
