cd $GRAFT_REPO_ROOT
run() { timeout 300 python bench.py --steps 40 --skip-matching --no-cpu-baseline > gpurun_out/sf.json 2>gpurun_out/sf.err; python -c "
import json; d=json.load(open('gpurun_out/sf.json')); print('$1', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']))"; }
ORB_FORK_EARLY=0 run "late fork, blur first, blur_ctas=8 (default)"
for bc in 7 6 5 4; do ORB_FORK_EARLY=0 ORB_BLUR_CTAS=$bc run "late fork, blur first, blur_ctas=$bc"; done
for bc in 8 6 5 4 3; do ORB_FORK_EARLY=3 ORB_BLUR_CTAS=$bc run "late fork, select first, blur_ctas=$bc"; done
