set -u
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_requests_srcunit_ltcfabric.sum
python tools/profile_target.py --scale 24 --prepared --no-extras --reps 1 > gpurun_out/r2_prof_plain24p.log 2>&1 &&
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_c3_prepared.csv \
    python tools/profile_target.py --scale 24 --prepared --no-extras --reps 1 > gpurun_out/r2_prof_ncu24p.log 2>&1
python tools/profile_target.py --scale 24 --no-extras --reps 1 > gpurun_out/r2_prof_plain24s.log 2>&1 &&
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_c3_stateless.csv \
    python tools/profile_target.py --scale 24 --no-extras --reps 1 > gpurun_out/r2_prof_ncu24s.log 2>&1
python tools/profile_target.py --scale 24 --prepared --no-extras --reps 1 --option edge_filter_mb=32 > gpurun_out/r2_prof_plain24f.log 2>&1 &&
ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_c3_hubfilter32.csv \
    python tools/profile_target.py --scale 24 --prepared --no-extras --reps 1 --option edge_filter_mb=32 > gpurun_out/r2_prof_ncu24f.log 2>&1
tail -n 2 gpurun_out/r2_prof_ncu24p.log; tail -n 2 gpurun_out/r2_prof_ncu24f.log
