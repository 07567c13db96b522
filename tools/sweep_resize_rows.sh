cd $GRAFT_REPO_ROOT
run() { timeout 300 python bench.py --steps 40 --skip-matching --no-cpu-baseline > gpurun_out/sr.json 2>gpurun_out/sr.err; python -c "
import json; d=json.load(open('gpurun_out/sr.json')); print('$1', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), round(d['roofline']['stage_ms_per_step']['k_resize(x7)'],4))"; }
run "rows=8 (default)"
for r in 16 12 24 4; do ORB_RESIZE_ROWS=$r run "rows=$r"; done
run "rows=8 (again)"
ORB_RESIZE_ROWS=16 timeout 600 python -m pytest tests/test_gpu_extract.py -x -q -m gpu 2>&1 | tail -3
