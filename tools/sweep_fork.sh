cd $GRAFT_REPO_ROOT
run() { timeout 300 python bench.py --steps 20 --skip-matching --no-cpu-baseline > gpurun_out/sf.json 2>gpurun_out/sf.err; python -c "
import json; d=json.load(open('gpurun_out/sf.json')); print('$1', round(d['value']), d['ms_per_step'])"; }
ORB_FORK_EARLY=0 run "late fork (baseline)"
for fc in 8 7 6 5; do for bc in 1 2; do ORB_FORK_EARLY=1 ORB_FAST_CTAS_FORK=$fc ORB_BLUR_CTAS=$bc run "early fast_ctas=$fc blur_ctas=$bc"; done; done
