#!/bin/bash
# ncu passes over one short bench run (B200_PROFILING.md recipe): launch list with durations, then a full capture of
# the path's kernels of one step.  Each ncu run follows a plain run of the same command that exited 0.
# The full capture skips the three warm-up steps (NCU_SKIP kernels of the path; 10 per step of the headline workload).
# TAG names the outputs (gpurun_out/${TAG}_launches.csv, ${TAG}_prof.ncu-rep); BENCH_ARGS selects the workload.
set -u
mkdir -p gpurun_out
TAG=${TAG:-r02}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { echo BUILD FAILED; tail gpurun_out/build.log; exit 1; }
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-stages --no-scale-sweep --no-gpu-baseline ${BENCH_ARGS:-}"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launches.log 2>&1
echo "launch list rc=$?"
if [ "${FULLCAP:-1}" = "1" ]; then
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"${NCU_KERNELS:-composite|preprocess|tile_|scan_block|coarse_scatter|emit|onesweep|histogram}" -s ${NCU_SKIP:-30} -c ${NCU_COUNT:-10} -f -o gpurun_out/${TAG}_prof $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
echo "full capture rc=$?"; tail -2 gpurun_out/${TAG}_ncu_full.log
fi
ls -la gpurun_out/ | grep ${TAG}
