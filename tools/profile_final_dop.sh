T=r2final
F="--no-cpu-baseline --no-other-workloads"
C5="python bench.py --workload config5 --steps 2 --warmup 1 $F --scale 0.05"
$C5 > gpurun_out/${T}_plain_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:dop853_kernel -s 3 -c 1 -f -o gpurun_out/${T}_prof_dop $C5 > gpurun_out/${T}_ncu_c5.log 2>&1
python tools/ncu_summary.py gpurun_out/${T}_prof_dop.ncu-rep gpurun_out/${T}_ncu_dop853_v2.md "ncu --set full: dop853_v2 ($T, last commit of round 2)" "$C5" > /dev/null
rm -f gpurun_out/${T}_prof_dop.ncu-rep
python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err
python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2>> gpurun_out/${T}_bench.err
