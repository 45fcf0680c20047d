#!/bin/bash
# Round profile: launch list of the bench command + one ncu --set full capture per integrator kernel.
# usage (GPU box): bash tools/profile_round.sh TAG
T=${1:-r2}
F="--no-cpu-baseline --no-other-workloads"
C4="python bench.py --steps 2 --warmup 1 $F"
C3="python bench.py --workload config3 --steps 2 --warmup 1 $F --scale 0.2"
C3K="python bench.py --workload config3k --steps 2 --warmup 1 $F --scale 0.2"
C5="python bench.py --workload config5 --steps 2 --warmup 1 $F --scale 0.05"
LM="python bench.py --workload config4_lm10 --steps 2 --warmup 1 $F --scale 0.3"
$C4 > gpurun_out/${T}_plain_c4.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches.csv $C4 > gpurun_out/${T}_ncu_launch.log 2>&1
$C4 > /dev/null 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:leapfrog_kernel -s 1 -c 1 -f -o gpurun_out/${T}_prof_lf $C4 > gpurun_out/${T}_ncu_lf.log 2>&1
$C3 > gpurun_out/${T}_plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:leapfrog_kernel -s 1 -c 1 -f -o gpurun_out/${T}_prof_store $C3 > gpurun_out/${T}_ncu_c3.log 2>&1
$C3K > gpurun_out/${T}_plain_c3k.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:leapfrog_kernel -s 1 -c 1 -f -o gpurun_out/${T}_prof_storek $C3K > gpurun_out/${T}_ncu_c3k.log 2>&1
$C5 > gpurun_out/${T}_plain_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dop853_kernel -s 3 -c 1 -f -o gpurun_out/${T}_prof_dop $C5 > gpurun_out/${T}_ncu_c5.log 2>&1
$LM > gpurun_out/${T}_plain_lm10.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:leapfrog_kernel -s 1 -c 1 -f -o gpurun_out/${T}_prof_lm10 $LM > gpurun_out/${T}_ncu_lm10.log 2>&1
# summaries and raw metric tables are made here, on the box: five reports exceed the 64 MiB that travel back
for k in lf:leapfrog_v2:"$C4" store:leapfrog_v2_store_all:"$C3" storek:leapfrog_v2_store_all_kicked:"$C3K" dop:dop853_v2:"$C5" lm10:leapfrog_lm10:"$LM"; do
  IFS=: read a b c <<< "$k"
  if [ -f gpurun_out/${T}_prof_$a.ncu-rep ]; then
    python tools/ncu_summary.py gpurun_out/${T}_prof_$a.ncu-rep gpurun_out/${T}_ncu_$b.md "ncu --set full: $b ($T)" "$c" > /dev/null
    ncu -i gpurun_out/${T}_prof_$a.ncu-rep --page raw --csv > gpurun_out/${T}_raw_$a.csv 2>/dev/null
    [ "$KEEP_REP" = "$a" ] || rm -f gpurun_out/${T}_prof_$a.ncu-rep
  fi
done
ls -la gpurun_out/${T}_*
