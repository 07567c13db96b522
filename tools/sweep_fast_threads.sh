cd $GRAFT_REPO_ROOT
cp orbslam_jpminipc_b200/liborb_b200.so /tmp/orig.so
trap 'cp /tmp/orig.so orbslam_jpminipc_b200/liborb_b200.so' EXIT     # round-1 script: swaps the product library; always restore it (newer A/B runs use tools/ab_lib.sh + ORB_B200_LIB instead)
for t in 128 96 64; do
  cp orbslam_jpminipc_b200/liborb_b200_$t.so orbslam_jpminipc_b200/liborb_b200.so
  timeout 300 python bench.py --steps 20 --skip-matching --no-cpu-baseline > gpurun_out/sw_$t.json 2>gpurun_out/sw.err
  python -c "
import json; d=json.load(open('gpurun_out/sw_$t.json')); print($t, round(d['value']), d['ms_per_step'], round(d['roofline']['stage_ms_per_step']['k_fast_nms'],3))"
done
cp /tmp/orig.so orbslam_jpminipc_b200/liborb_b200.so
