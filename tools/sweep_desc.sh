cd $GRAFT_REPO_ROOT
run() { python bench.py --steps 40 --skip-matching --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('$1', round(d['value']), round(d['roofline']['stage_ms_per_step']['k_describe'],4))"; }
run "6 CTAs/SM (40 regs)"
for n in 7 8; do ORB_B200_LIB=$PWD/orbslam_jpminipc_b200/variant_desc_$n.so run "$n CTAs/SM"; done
