cd $GRAFT_REPO_ROOT
run() { python bench.py --steps 40 --skip-matching --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('$1', round(d['value']), round(d['roofline']['stage_ms_per_step']['k_fast_nms'],4))"; }
run "FAST_CTAS=8 (default)"
ORB_B200_LIB=$PWD/orbslam_jpminipc_b200/variant_fast6.so run "FAST_CTAS=6"
run "FAST_CTAS=8 (default, again)"
