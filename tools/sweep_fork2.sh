cd $GRAFT_REPO_ROOT
run() { timeout 300 python bench.py --steps 40 --skip-matching --no-cpu-baseline > gpurun_out/sf.json 2>gpurun_out/sf.err; python -c "
import json; d=json.load(open('gpurun_out/sf.json')); print('$1', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']))"; }
ORB_FORK_EARLY=0 run "late fork (default)"
ORB_FORK_EARLY=2 ORB_BLUR_CTAS=4 run "fork behind FAST blur_ctas=4"
ORB_FORK_EARLY=2 ORB_BLUR_CTAS=8 run "fork behind FAST blur_ctas=8"
ORB_FORK_EARLY=2 ORB_BLUR_CTAS=2 run "fork behind FAST blur_ctas=2"
for fc in 6 5 4; do for bc in 1 2; do ORB_FORK_EARLY=1 ORB_FAST_CTAS_FORK=$fc ORB_BLUR_CTAS=$bc run "beside FAST fast_ctas=$fc blur_ctas=$bc"; done; done
ORB_FORK_EARLY=0 run "late fork (default, again)"
