#!/bin/bash
# Host-mode chunking of batches below four waves: the default cut against ONE chunk (CW_HOST_CHUNK >= batch), GPU box
run() {  # workload scale
  for c in default single; do
    if [ $c = single ]; then export CW_HOST_CHUNK=100000000; else unset CW_HOST_CHUNK; fi
    python bench.py --workload $1 --scale $2 --steps 5 --warmup 3 --no-cpu-baseline --no-other-workloads 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$1 x$2 (%d orbits) $c: value %.4g  e2e %.4g (%.2f ms)' % (d['run']['n_orbits_total'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step']))
"
  done
}
run config2 1.0; run config2 2.0; run config2 3.0
run config4 0.02; run config4 0.03; run config4 0.045; run config4 0.06
