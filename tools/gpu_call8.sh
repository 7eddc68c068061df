mkdir -p gpurun_out
for v in 1 0; do
  LRF_NVCC_EXTRA="-DLRF_STEAL=$v" python -c "from localrf_b200 import _lib; _lib.build(force=True)" || continue
  echo "=== LRF_STEAL=$v"
  timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_local.py tests/test_gpu_cabi_direct.py tests/test_gpu_fullsize.py tests/test_gpu_vs_reference.py tests/test_gpu_pipeline.py -q -x -p no:cacheprovider > gpurun_out/r2_steal${v}_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_steal${v}_tests.log
  for w in cfg2 distB cfg3 cfg5; do
    timeout 600 python bench.py --steps 40 --warmup 5 --workload $w --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_steal${v}_$w.json 2>gpurun_out/r2_steal${v}_$w.err || { echo "bench $w failed"; tail -3 gpurun_out/r2_steal${v}_$w.err; continue; }
    python -c "
import json; d=json.load(open('gpurun_out/r2_steal${v}_$w.json')); print('$w value %.2fM ms %.4f e2e %.2fM sync %.2fM frame %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['sync_per_step_value']/1e6, d['frame_api'] and round(d['frame_api']['rays_per_s']/1e6,2)))"
  done
  timeout 300 python tools/small_batch.py
done
