cd $GRAFT_REPO_ROOT
for v in "$@"; do
  if [ "$v" = default ]; then unset ORB_B200_LIB; else export ORB_B200_LIB=$PWD/orbslam_jpminipc_b200/liborb_b200_$v.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --skip-matching --no-cpu-baseline --quick > gpurun_out/ab_x_$v.json 2> gpurun_out/ab_x_$v.err || { echo "$v failed"; tail -5 gpurun_out/ab_x_$v.err; continue; }
  python - "$v" gpurun_out/ab_x_$v.json <<'P'
import json, sys
d = json.loads(open(sys.argv[2]).read().strip().splitlines()[-1])
st = d["roofline"]["stage_ms_per_launch"]
print(sys.argv[1], "value", round(d["value"]), "ms/launch", round(d["ms_per_step"] / d["step"]["launches"], 4), {k: round(v, 4) for k, v in st.items()}, "verified" if d.get("verified") else "NOT VERIFIED")
P
done
