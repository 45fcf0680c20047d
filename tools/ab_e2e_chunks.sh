#!/bin/bash
# e2e with and without wave-quantised chunks (same library, env switch)
for W in ${@:-config4 config2 config4_lm10 config4_evolving}; do
  for V in wave old; do
    if [ $V = old ]; then export CW_NO_WAVE_CHUNKS=1; else unset CW_NO_WAVE_CHUNKS; fi
    python bench.py --workload $W --steps 3 --warmup 2 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$W', '$V', 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'e2e_ms %.3f' % d['e2e']['ms_per_step'])
"
  done
done
