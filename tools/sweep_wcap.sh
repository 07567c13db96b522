cd $GRAFT_REPO_ROOT
run() { python bench.py --steps 40 --skip-matching --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.readlines()[-1]); print('$1', round(d['value']), round(d['roofline']['stage_ms_per_step']['k_select'],4), round(d['single_frame_latency']['graph_replay']['median_ms'],4), round(d['config0_640x480']['value']))"; }
for wc in 512 384 256; do ORB_B200_LIB=$PWD/orbslam_jpminipc_b200/variant_wcap$wc.so run "SEL_WCAP=$wc"; done
