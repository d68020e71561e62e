set -x
O=gpurun_out
for k in bin_xform bin_tri raster_binned; do
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s 1 -c 1 -o $O/r02z_cfg3_$k -f python profiles/staged_workloads.py 3 1024 2 > /dev/null 2>&1
  ncu --set full --clock-control none --import-source on -k "regex:$k" -s 1 -c 1 -o $O/r02z_cfg5_$k -f python profiles/staged_workloads.py 5 512 2 > /dev/null 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02z_cfg3_launches.csv python profiles/staged_workloads.py 3 1024 3 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r02z_cfg5_launches.csv python profiles/staged_workloads.py 5 1024 2 > /dev/null 2>&1
python profiles/staged_workloads.py 3 1024 10 > $O/r02z_cfg_times.log 2>&1
python profiles/staged_workloads.py 5 1024 3 >> $O/r02z_cfg_times.log 2>&1
PBR_B200_LARGE=staged python profiles/staged_workloads.py 3 1024 10 >> $O/r02z_cfg_times.log 2>&1
PBR_B200_LARGE=staged python profiles/staged_workloads.py 5 1024 3 >> $O/r02z_cfg_times.log 2>&1
timeout 300 compute-sanitizer --tool memcheck python profiles/staged_workloads.py 3 16 1 > $O/r02z_memcheck_binned.log 2>&1
timeout 300 compute-sanitizer --tool racecheck python profiles/staged_workloads.py 5 4 1 > $O/r02z_racecheck_binned.log 2>&1
grep "^config" $O/r02z_cfg_times.log; tail -n 1 $O/r02z_memcheck_binned.log $O/r02z_racecheck_binned.log
