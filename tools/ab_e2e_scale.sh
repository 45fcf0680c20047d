#!/bin/bash
# per-rank batch sizes of the strong-scaling run (N = 4, 8) on one GPU: wave-quantised chunks against the old cut
for S in 0.25 0.125; do
  for V in wave old; do
    if [ $V = old ]; then export CW_NO_WAVE_CHUNKS=1; else unset CW_NO_WAVE_CHUNKS; fi
    python bench.py --scale $S --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('scale $S', '$V', 'value %.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], 'e2e_ms %.3f' % d['e2e']['ms_per_step'])
"
  done
done
