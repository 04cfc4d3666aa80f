#!/bin/bash
# Developer experiment: compositing staging-batch sizes (LGM_FWD_BATCH / LGM_BWD_BATCH)
for fb in ${BATCHES:-256 512 768 1024}; do
    LGM_FWD_BATCH=$fb LGM_BWD_BATCH=$fb python bench.py --steps 5 --no-e2e --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_ms']
print('batch', $fb, 'fwd %.2f bwd %.2f step %.2f' % (s['composite_fwd'], s['composite_bwd'], d['ms_per_step']))"
done
