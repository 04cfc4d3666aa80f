#!/bin/bash
# Developer experiment: resident CTAs per SM (launch bounds) x staging batch of the two-pixel compositing kernels.
# CFGS="occ:fwdbatch:bwdbatch ..."
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for cfg in ${CFGS:-1006:256:384 1008:256:384 1010:256:384 1006:256:256 1008:256:256}; do
IFS=: read occ fb bb <<< "$cfg"
LGM_C2_OCC=$occ LGM_FWD_BATCH=$fb LGM_BWD_BATCH=$bb python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-scale-sweep --no-gpu-baseline ${BENCH_ARGS:-} 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
s = d['stages_ms']
print('occ $occ batch $fb/$bb: step %.3f ms  fwd %.3f  bwd %.3f' % (d['ms_per_step'], s['composite_fwd'], s['composite_bwd']))"
done
