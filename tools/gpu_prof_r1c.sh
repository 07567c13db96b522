set -x
cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 2 --warmup 3 --skip-matching --no-cpu-baseline --batch 64"
$CMD > gpurun_out/plain_c.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_' -s 45 -c 15 -o gpurun_out/prof_all_r1c $CMD > gpurun_out/ncu_c.log 2>&1
tail -3 gpurun_out/ncu_c.log
ls -la gpurun_out/*.ncu-rep
