#!/bin/bash
# One profiling pass on the GPU box:   gpurun --timeout 900 -- 'bash tools/gpu_prof.sh r2a [full] [extra bench flags]'
#   launch list  (ncu --metrics gpu__time_duration.sum)   -> gpurun_out/launches_<tag>.csv
#   full capture (ncu --set full --import-source on, only with the word "full") -> gpurun_out/prof_<tag>.ncu-rep
# Each ncu pass runs only after the same command has exited 0 without ncu.  Summaries for profiles/ are made here afterwards
# (tools/ncu_summary.py, tools/launch_shares.py).
TAG=${1:?tag}; shift
FULL=0; if [ "$1" = full ]; then FULL=1; shift; fi
cd "$GRAFT_REPO_ROOT"
CMD="python bench.py --steps 2 --warmup 3 --skip-matching --no-cpu-baseline --no-verify --batch 64 --step-launches 1 $*"
timeout 120 $CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu1_$TAG.log 2>&1
echo "launch list rc=$?"
if [ $FULL = 1 ]; then
  timeout 120 $CMD > gpurun_out/plain2_$TAG.log 2>&1 &&
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_' -s 60 -c 16 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu2_$TAG.log 2>&1
  echo "full capture rc=$?"; ls -la gpurun_out/prof_$TAG.ncu-rep
fi
