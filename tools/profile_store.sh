#!/bin/bash
T=${1:-r1f3}
C3="python bench.py --workload config3 --steps 2 --warmup 1 --no-cpu-baseline --scale 0.2"
C3K="python bench.py --workload config3k --steps 2 --warmup 1 --no-cpu-baseline --scale 0.2"
$C3 > gpurun_out/${T}_plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:leapfrog_kernel -s 1 -c 1 -f -o gpurun_out/${T}_prof_store $C3 > gpurun_out/${T}_ncu_c3.log 2>&1
$C3K > gpurun_out/${T}_plain_c3k.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:leapfrog_kernel -s 1 -c 1 -f -o gpurun_out/${T}_prof_storek $C3K > gpurun_out/${T}_ncu_c3k.log 2>&1
ls -la gpurun_out/${T}_prof_*
