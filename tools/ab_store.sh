#!/bin/bash
# A/B the in-tree library against build_variants/*.so on the store_all workloads (GPU box)
for W in config3 config3k; do
for lib in cogsworth_b200/libcogsworth_b200.so $(ls build_variants/*.so 2>/dev/null); do
  CW_DEBUG=1 COGSWORTH_B200_LIB=$PWD/$lib python bench.py --workload $W --steps 3 --warmup 2 --no-cpu-baseline --scale ${SCALE:-0.3} 2>gpurun_out/ab_err.txt | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('$W', '$lib', 'value %.4g' % d['value'], 'kernel_ms %.3f' % r['kernel_ms'], 'achieved %.0f' % r['achieved'], 'frac %.3f' % r['frac'])
"
  grep -m2 "\[cw\]" gpurun_out/ab_err.txt
done
done
