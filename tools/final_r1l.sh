cd $GRAFT_REPO_ROOT
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/gputests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.log 2>gpurun_out/bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py > gpurun_out/bench1.log 2>gpurun_out/bench1.err; echo "bench rc=$?"
bash tools/gpu_prof_r1l.sh > gpurun_out/prof_r1l.log 2>&1; echo "prof rc=$?"
