#!/bin/bash
# One GPU-box pass: sort tests first (short timeout: a hung look-back must not eat the budget), then the parity
# suite, smoke and a short bench.  Everything lands in gpurun_out/.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail -20 gpurun_out/build.log; exit 1; }
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short -k "onesweep" > gpurun_out/pytest_sort.log 2>&1
rc=$?; echo "sort tests rc=$rc"; tail -5 gpurun_out/pytest_sort.log
if [ $rc -eq 124 ]; then echo "sort tests TIMED OUT - stopping"; exit 2; fi
timeout 900 python -m pytest tests -m gpu -q --tb=short --maxfail=12 -k "not onesweep" > gpurun_out/pytest_gpu.log 2>&1
echo "parity tests rc=$?"; tail -40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
