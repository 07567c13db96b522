#!/bin/bash
# A/B of one environment knob on the GPU box:  gpurun -- 'bash tools/ab_env.sh TAG VAR v1 v2 ...'
# runs the quick headline bench (640x480, verification on) once per value and prints value, step and stage times
TAG=${1:?tag}; VAR=${2:?var}; shift 2
cd "$GRAFT_REPO_ROOT"
for v in "$@"; do
  env $VAR=$v timeout 300 python bench.py --steps 5 --warmup 3 --skip-matching --no-cpu-baseline --quick > gpurun_out/ab_${TAG}_$v.json 2> gpurun_out/ab_${TAG}_$v.err || { echo "$VAR=$v failed"; tail -5 gpurun_out/ab_${TAG}_$v.err; continue; }
  python - "$VAR=$v" gpurun_out/ab_${TAG}_$v.json <<'P'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
st = d["roofline"]["stage_ms_per_launch"]
print(sys.argv[1], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/launch", round(d["ms_per_step"] / d["step"]["launches"], 4),
      {k: round(v, 4) for k, v in st.items()}, "verified" if d.get("verified") else "NOT VERIFIED")
P
done
