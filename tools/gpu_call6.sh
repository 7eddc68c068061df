mkdir -p gpurun_out
for t in 640 768 512; do
  LRF_NVCC_EXTRA="-DLRF_THREADS=$t" python -c "from localrf_b200 import _lib; _lib.build(force=True)" || continue
  echo "=== THREADS=$t"
  python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider 2>&1 | tail -2
  for w in cfg2 distB cfg5; do
    python bench.py --steps 40 --warmup 5 --workload $w --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_thr${t}_$w.json 2>/dev/null
    python -c "
import json; d=json.load(open('gpurun_out/r2_thr${t}_$w.json')); print('$w value %.2fM ms %.4f e2e %.2fM sync %.2fM frame %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['sync_per_step_value']/1e6, d['frame_api'] and round(d['frame_api']['rays_per_s']/1e6,2)))"
  done
done
