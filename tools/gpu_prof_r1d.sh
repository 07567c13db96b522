set -x
cd $GRAFT_REPO_ROOT
CMD="python bench.py --steps 2 --warmup 3 --skip-matching --no-cpu-baseline --batch 64"
$CMD > gpurun_out/plain_d.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1d.csv $CMD > gpurun_out/ncu_d1.log 2>&1
$CMD > gpurun_out/plain_d2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_' -s 60 -c 15 -o gpurun_out/prof_all_r1d $CMD > gpurun_out/ncu_d2.log 2>&1
CMD2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --batch 32 --db-rows 2000000"
$CMD2 > gpurun_out/plain_d3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_knn2$' -s 3 -c 2 -o gpurun_out/prof_knn_r1d $CMD2 > gpurun_out/ncu_d3.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r1d.csv
python -m pytest tests -m gpu -q 2>&1 | tail -2
