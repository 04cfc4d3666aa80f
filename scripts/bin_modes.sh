#!/bin/bash
# Developer experiment: the three binning paths of lgm_forward_bin on the three regimes
for a in "" "--kind init" "--workload scale_sweep"; do for m in direct onesweep hybrid; do
LGM_BIN_MODE=$m python bench.py $a --steps 3 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$m', '$a', 'step', round(d['ms_per_step'],2), 'bin', round(d['stages_ms']['bin'],2))"
done; done
