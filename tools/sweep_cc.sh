cd $GRAFT_REPO_ROOT
run() { python bench.py --steps 40 --skip-matching --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('$1', round(d['value']), round(d['roofline']['stage_ms_per_step']['k_cell_compact'],4))"; }
run "default (47 regs)"
for n in 6 8; do ORB_B200_LIB=$PWD/orbslam_jpminipc_b200/variant_cc_$n.so run "$n CTAs/SM"; done
