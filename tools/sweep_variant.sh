# A/B of a k_fast_nms build variant (e.g. -DORB_NMS_LOBYTE) WITHOUT touching the product library: the variant is a second .so selected
# through ORB_B200_LIB (orbslam_jpminipc_b200/_lib.py).
# Here (build container):   make -C orbslam_jpminipc_b200/csrc -s OUT=../liborb_b200_variant.so EXTRA=-DORB_NMS_LOBYTE
# then:                     gpurun --timeout 600 -- 'bash tools/sweep_variant.sh > gpurun_out/sweep_variant.log 2>&1; cat gpurun_out/sweep_variant.log'
cd "$GRAFT_REPO_ROOT"
VAR=$PWD/orbslam_jpminipc_b200/liborb_b200_variant.so
run() { timeout 300 python bench.py --steps 10 --skip-matching --no-cpu-baseline > gpurun_out/sr.json 2>gpurun_out/sr.err; python -c "
import json; d=json.load(open('gpurun_out/sr.json')); s=d['roofline']['stage_ms_per_step']; print('$1', round(d['value']), round(d['ms_per_step'],4), round(d['e2e']['value']), {k: round(v,4) for k,v in s.items()})" || tail -3 gpurun_out/sr.err; }
run "default"
ORB_B200_LIB=$VAR timeout 900 python -m pytest tests/test_gpu_extract.py tests/test_gpu_vs_ref.py tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -2
ORB_B200_LIB=$VAR run "variant"
run "default (again)"
