cd $GRAFT_REPO_ROOT
run() { timeout 300 python bench.py --steps 40 --skip-matching --no-cpu-baseline > gpurun_out/sf.json 2>gpurun_out/sf.err; python -c "
import json; d=json.load(open('gpurun_out/sf.json')); print('$1', round(d['ms_per_step'],4))" || tail -3 gpurun_out/sf.err; }
run "full pipeline"
ORB_DEBUG_SKIP=1 run "no blur"
ORB_DEBUG_SKIP=2 run "no selection (describe sees stale lists)"
ORB_DEBUG_SKIP=3 run "no blur, no selection"
ORB_DEBUG_SKIP=4 run "no describe"
ORB_DEBUG_SKIP=7 run "no blur, selection, describe"
