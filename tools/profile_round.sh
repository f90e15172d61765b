#!/bin/bash
# The ncu passes behind profiles/r02_*: launch lists of one kept-graph and one stateless c3 call (time, DRAM bytes,
# L2->fabric requests per launch) and one --set full capture of the walk kernel and the table builders at scale 23.
# Run under gpurun from the repository root; every ncu pass follows a plain run of the same command that exited 0.
set -u
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_ltcfabric.sum
python tools/profile_target.py --scale 24 --prepared --no-extras --reps 1 > gpurun_out/r2_prof_plain24p.log 2>&1 &&
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_c3_prepared.csv \
    python tools/profile_target.py --scale 24 --prepared --no-extras --reps 1 > gpurun_out/r2_prof_ncu24p.log 2>&1
python tools/profile_target.py --scale 24 --no-extras --reps 1 > gpurun_out/r2_prof_plain24s.log 2>&1 &&
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_c3_stateless.csv \
    python tools/profile_target.py --scale 24 --no-extras --reps 1 > gpurun_out/r2_prof_ncu24s.log 2>&1
python tools/profile_target.py --scale 23 --prepared --no-extras --reps 1 > gpurun_out/r2_prof_plain23.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"node2vec_walk|build_tiled|build_hub|edge_bloom" -c 6 \
    -o gpurun_out/r2_prof_s23 -f python tools/profile_target.py --scale 23 --prepared --no-extras --reps 1 > gpurun_out/r2_prof_ncu23.log 2>&1
ls -la gpurun_out/r2_prof_s23.ncu-rep
