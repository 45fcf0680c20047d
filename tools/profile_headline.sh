#!/bin/bash
T=${1:-r1f4}
C4="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
$C4 > gpurun_out/${T}_plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $C4 > gpurun_out/${T}_ncu_launch.log 2>&1
$C4 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:leapfrog_kernel -s 1 -c 1 -f -o gpurun_out/${T}_prof_lf $C4 > gpurun_out/${T}_ncu_lf.log 2>&1
ls -la gpurun_out/${T}_*
