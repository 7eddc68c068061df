mkdir -p gpurun_out
for v in "512 8 0" "640 8 0" "640 4 0" "768 8 0" "512 4 200" "512 4 0"; do
  set -- $v
  LRF_NVCC_EXTRA="-DLRF_THREADS=$1 -DLRF_CONS_WARPS=$2 -DLRF_SPIN_NS=$3" python -c "from localrf_b200 import _lib; _lib.build(force=True)" || continue
  echo "=== THREADS=$1 CONS_WARPS=$2 SPIN_NS=$3"
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_local.py -q -x -p no:cacheprovider 2>&1 | tail -1
  for w in cfg2 distB cfg5; do
    python bench.py --steps 40 --warmup 5 --workload $w --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_thr$1_$2_$3_$w.json 2>/dev/null
    python -c "
import json; d=json.load(open('gpurun_out/r2_thr$1_$2_$3_$w.json')); print('$w value %.2fM ms %.4f e2e %.2fM sync %.2fM frame %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['sync_per_step_value']/1e6, d['frame_api'] and round(d['frame_api']['rays_per_s']/1e6,2)))"
  done
done
