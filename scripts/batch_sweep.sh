#!/bin/bash
# Developer experiment: staging batch of the compositing kernels (LGM_FWD_BATCH / LGM_BWD_BATCH) on the headline workload.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
for b in ${BATCHES:-256 384 512 640 768 1024}; do
LGM_FWD_BATCH=$b LGM_BWD_BATCH=$b python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-scale-sweep --no-gpu-baseline ${BENCH_ARGS:-} 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read().strip().splitlines()[-1])
s = d['stages_ms']
print('batch $b: step %.3f ms  fwd %.3f  bwd %.3f' % (d['ms_per_step'], s['composite_fwd'], s['composite_bwd']))"
done
