set -x
cd $GRAFT_REPO_ROOT
for ch in 16 32 64 128 256; do python bench.py --steps 20 --warmup 3 --skip-matching --no-cpu-baseline --chunk $ch > gpurun_out/sweep_chunk$ch.json 2>gpurun_out/sweep.err || tail -3 gpurun_out/sweep.err; done
python bench.py --steps 2 --warmup 3 --skip-matching --no-cpu-baseline --batch 64 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 2 --warmup 3 --skip-matching --no-cpu-baseline --batch 64 > gpurun_out/ncu1.log 2>&1
python bench.py --steps 2 --warmup 3 --skip-matching --no-cpu-baseline --batch 64 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_fast_nms -s 3 -c 1 -o gpurun_out/prof_fast_r1a python bench.py --steps 2 --warmup 3 --skip-matching --no-cpu-baseline --batch 64 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
