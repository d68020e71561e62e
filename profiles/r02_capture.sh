#!/bin/bash
# One gpurun call: every raw profile of round 2 (final build).  Summaries are made locally by profiles/r02_summarise.py.
set -x
O=gpurun_out
python bench.py --steps 200 --warmup 20 > $O/r02s_bench.json 2> $O/r02s_bench.err
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > $O/r02s_bench20.json 2>> $O/r02s_bench.err
ncu --set full --clock-control none --import-source on -k regex:raster_warp -s 6 -c 1 -o $O/r02z_warp -f python bench.py --no-cpu-baseline --no-extras --no-verify --steps 16 --warmup 3 > $O/r02z_warp.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $O/r02z_launches.csv python bench.py --no-cpu-baseline --no-extras --no-verify --steps 64 --warmup 3 > /dev/null 2>&1
ncu --replay-mode range --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none python profiles/traffic_range.py 16 > $O/r02z_traffic.log 2>&1
for k in bin_xform bin_tri "bin_blocks_kernel<0>" "bin_blocks_kernel<1>" raster_binned; do
  n=$(echo $k | tr -d '<>' )
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s 1 -c 1 -o $O/r02z_cfg3_$n -f python profiles/staged_workloads.py 3 1024 2 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s 1 -c 1 -o $O/r02z_cfg5_$n -f python profiles/staged_workloads.py 5 512 2 > /dev/null 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02z_cfg3_launches.csv python profiles/staged_workloads.py 3 1024 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02z_cfg5_launches.csv python profiles/staged_workloads.py 5 1024 2 > /dev/null 2>&1
python profiles/staged_workloads.py 3 1024 10 > $O/r02z_cfg_times.log 2>&1
python profiles/staged_workloads.py 5 1024 3 >> $O/r02z_cfg_times.log 2>&1
PBR_B200_LARGE=staged python profiles/staged_workloads.py 3 1024 10 >> $O/r02z_cfg_times.log 2>&1
PBR_B200_LARGE=staged python profiles/staged_workloads.py 5 1024 3 >> $O/r02z_cfg_times.log 2>&1
timeout 300 compute-sanitizer --tool racecheck python -c "import __graft_entry__ as g; g.smoke()" > $O/r02z_racecheck.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python -c "import __graft_entry__ as g; g.smoke()" > $O/r02z_memcheck.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python profiles/staged_workloads.py 3 16 1 > $O/r02z_memcheck_binned.log 2>&1
timeout 300 compute-sanitizer --tool racecheck python profiles/staged_workloads.py 5 4 1 > $O/r02z_racecheck_binned.log 2>&1
for f in $O/r02z_racecheck.log $O/r02z_memcheck.log $O/r02z_memcheck_binned.log $O/r02z_racecheck_binned.log; do tail -n 3 $f; done
