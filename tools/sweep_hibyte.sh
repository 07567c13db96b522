# A/B of the ORB_NMS_HIBYTE variant of k_fast_nms ("Next candidates" (a)).
# Here (build container):   make -C orbslam_jpminipc_b200/csrc -s OUT=../liborb_b200_hibyte.so EXTRA=-DORB_NMS_HIBYTE && make -C orbslam_jpminipc_b200/csrc -s
# then:                     gpurun --timeout 600 -- 'bash tools/sweep_hibyte.sh > gpurun_out/sweep_hibyte.log 2>&1; cat gpurun_out/sweep_hibyte.log'
cd $GRAFT_REPO_ROOT
cp orbslam_jpminipc_b200/liborb_b200.so /tmp/orig.so
run() { timeout 300 python bench.py --steps 40 --skip-matching --no-cpu-baseline > gpurun_out/sr.json 2>gpurun_out/sr.err; python -c "
import json; d=json.load(open('gpurun_out/sr.json')); s=d['roofline']['stage_ms_per_step']; print('$1', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), round(s['k_fast_nms'],4))"; }
run "default"
cp orbslam_jpminipc_b200/liborb_b200_hibyte.so orbslam_jpminipc_b200/liborb_b200.so
timeout 600 python -m pytest tests/test_gpu_extract.py tests/test_gpu_vs_ref.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
run "ORB_NMS_HIBYTE"
cp /tmp/orig.so orbslam_jpminipc_b200/liborb_b200.so
run "default (again)"
