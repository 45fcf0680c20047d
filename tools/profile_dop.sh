#!/bin/bash
T=${1:-r1z}
C5="python bench.py --workload config5 --steps 2 --warmup 1 --no-cpu-baseline --scale 0.05"
C5D="python bench.py --workload config5_dense --steps 2 --warmup 1 --no-cpu-baseline --scale 0.05"
$C5 > gpurun_out/${T}_plain_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dop853_kernel -s 3 -c 1 -f -o gpurun_out/${T}_prof_dop $C5 > gpurun_out/${T}_ncu_c5.log 2>&1
$C5D > gpurun_out/${T}_plain_c5d.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dop853_kernel -s 3 -c 1 -f -o gpurun_out/${T}_prof_dopd $C5D > gpurun_out/${T}_ncu_c5d.log 2>&1
ls -la gpurun_out/${T}_prof_dop*
