set -x
cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 2 --warmup 3 --skip-matching --no-cpu-baseline --batch 64"
$CMD > gpurun_out/plain_l.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1l.csv $CMD > gpurun_out/ncu_l1.log 2>&1
$CMD > gpurun_out/plain_l2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_' -s 60 -c 15 -o gpurun_out/prof_all_r1l $CMD > gpurun_out/ncu_l2.log 2>&1
ls -la gpurun_out/prof_all_r1l.ncu-rep gpurun_out/launches_r1l.csv
