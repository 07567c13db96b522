# A/B of run-time switches of the library (environment variables read at orb_create), e.g.
#   gpurun --timeout 600 -- 'bash tools/sweep_env.sh ORB_PYR_FUSED 1 0'
cd "$GRAFT_REPO_ROOT"
VAR=$1; shift
for v in "$@"; do
  env $VAR=$v timeout 200 python bench.py --steps 10 --quick --skip-matching --no-cpu-baseline > gpurun_out/se.json 2>gpurun_out/se.err && python -c "
import json; d=json.load(open('gpurun_out/se.json')); s=d['roofline']['stage_ms_per_launch']; print('$VAR=$v', round(d['value']), round(d['e2e']['value']), {k: round(x,4) for k,x in s.items()}, d['verified'] and 'verified')" || tail -3 gpurun_out/se.err
done
