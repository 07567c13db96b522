cd $GRAFT_REPO_ROOT
cp orbslam_jpminipc_b200/liborb_b200.so /tmp/orig.so
trap 'cp /tmp/orig.so orbslam_jpminipc_b200/liborb_b200.so' EXIT     # round-1 script: swaps the product library; always restore it (newer A/B runs use tools/ab_lib.sh + ORB_B200_LIB instead)
run() { timeout 300 python bench.py --steps 40 --skip-matching --no-cpu-baseline > gpurun_out/sr.json 2>gpurun_out/sr.err; python -c "
import json; d=json.load(open('gpurun_out/sr.json')); s=d['roofline']['stage_ms_per_step']; print('$1', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), round(s['k_fast_nms'],4))"; }
run "64x128 (default)"
for t in 128x64 64x64 128x32; do
  cp orbslam_jpminipc_b200/liborb_b200_t$t.so orbslam_jpminipc_b200/liborb_b200.so
  timeout 300 python -m pytest tests/test_gpu_extract.py -m gpu -q -x 2>&1 | tail -1
  run $t
done
cp /tmp/orig.so orbslam_jpminipc_b200/liborb_b200.so
