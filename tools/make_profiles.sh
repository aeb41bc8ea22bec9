#!/bin/bash
# GPU box: the launch list and one full capture of the dominant kernel, both from the bench command itself
# (each ncu pass only after the same command has exited 0 without ncu).  Outputs land in gpurun_out/.
CMD="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-workloads"
$CMD > gpurun_out/r02_plain.log 2> gpurun_out/r02_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_match_run -s 2 -c 1 -o gpurun_out/prof_r02 $CMD > gpurun_out/r02_ncu_full.log 2>&1
tail -2 gpurun_out/r02_ncu_full.log
