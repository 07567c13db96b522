cd $GRAFT_REPO_ROOT
run() { timeout 300 python bench.py --steps 40 --skip-matching --no-cpu-baseline > gpurun_out/sf.json 2>gpurun_out/sf.err; python -c "
import json; d=json.load(open('gpurun_out/sf.json')); print('$1', round(d['value']), round(d['roofline']['stage_ms_per_step']['k_select'],4), d['single_frame_latency']['graph_replay']['median_ms'])" || tail -3 gpurun_out/sf.err; }
run "warps=8 serial_below=8 (built default)"
for v in 4_8 2_8; do ORB_B200_LIB=$PWD/orbslam_jpminipc_b200/variant_sel_$v.so run "variant $v"; done
