cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_vs_ref.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -3
run() { timeout 300 python bench.py --steps 40 --skip-matching --no-cpu-baseline > gpurun_out/sr.json 2>gpurun_out/sr.err; python -c "
import json; d=json.load(open('gpurun_out/sr.json')); s=d['roofline']['stage_ms_per_step']; print('$1', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), {k:round(v,4) for k,v in s.items()})"; }
run "run 1"
run "run 2"
