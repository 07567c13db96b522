cd $GRAFT_REPO_ROOT
for ch in 32 64 96 128; do python bench.py --steps 40 --skip-matching --no-cpu-baseline --e2e-chunk $ch > gpurun_out/sf.json 2>gpurun_out/sf.err; python -c "
import json; d=json.load(open('gpurun_out/sf.json')); print('e2e_chunk $ch sync', round(d['e2e']['synchronous_call']['value']), 'stream', round(d['e2e']['value']))"; done
