cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 2 --warmup 3 --skip-matching --no-cpu-baseline --batch 64"
timeout 40 $CMD > gpurun_out/plain_m.log 2>&1 &&
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1m.csv $CMD > gpurun_out/ncu_m1.log 2>&1
echo rc=$?
