cd $GRAFT_REPO_ROOT
cp orbslam_jpminipc_b200/liborb_b200.so /tmp/orig.so
trap 'cp /tmp/orig.so orbslam_jpminipc_b200/liborb_b200.so' EXIT     # round-1 script: swaps the product library; always restore it (newer A/B runs use tools/ab_lib.sh + ORB_B200_LIB instead)
for t in 64x64 128x64 64x128 64x96; do
  cp orbslam_jpminipc_b200/liborb_b200_t$t.so orbslam_jpminipc_b200/liborb_b200.so
  timeout 200 python -m pytest tests/test_gpu_extract.py -m gpu -q -x 2>&1 | tail -1
  timeout 300 python bench.py --steps 20 --skip-matching --no-cpu-baseline > gpurun_out/st_$t.json 2>gpurun_out/st.err
  python -c "
import json; d=json.load(open('gpurun_out/st_$t.json')); print('$t', round(d['value']), d['ms_per_step'], round(d['roofline']['stage_ms_per_step']['k_fast_nms'],3))"
done
cp /tmp/orig.so orbslam_jpminipc_b200/liborb_b200.so
