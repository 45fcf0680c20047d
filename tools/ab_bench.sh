#!/bin/bash
# A/B the in-tree library against the builds in build_variants/ on one workload (GPU box).
W=${1:-config4}; shift
for lib in cogsworth_b200/libcogsworth_b200.so $(ls build_variants/*.so 2>/dev/null); do
  COGSWORTH_B200_LIB=$PWD/$lib python bench.py --workload $W --steps 3 --warmup 2 --no-cpu-baseline --no-other-workloads "$@" 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
r = d['roofline']
print('$lib', 'value %.4g' % d['value'], 'e2e', d['e2e'] and '%.4g' % d['e2e']['value'], 'kernel_ms %.3f' % r['kernel_ms'], 'frac %.3f' % r['frac'])
"
done
