#!/bin/bash
# compute-sanitizer over tools/gpu_sanitize.py (GPU box): memcheck, racecheck, synccheck, initcheck; summaries into gpurun_out/
T=${1:-r2}
for tool in memcheck racecheck synccheck initcheck; do
  for mode in "" "--ragged"; do
    tag=${T}_san_${tool}${mode:+_ragged}
    timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/gpu_sanitize.py $mode > gpurun_out/${tag}.log 2>&1
    echo "== $tool $mode: exit $? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' gpurun_out/${tag}.log | tail -1)  [$(grep -c '^done' gpurun_out/${tag}.log) done]"
  done
done
