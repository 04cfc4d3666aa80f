#!/bin/bash
# ncu passes over one short bench run (B200_PROFILING.md recipe): launch list with durations, then a full capture of
# the path's kernels of one step.  Each ncu run follows a plain run of the same command that exited 0.
set -u
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-stages ${BENCH_ARGS:-}"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-onesweep|composite|preprocess|emit|histogram|tile_ranges|scan_block}" -s ${NCU_SKIP:-0} -c ${NCU_COUNT:-16} -f -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/
