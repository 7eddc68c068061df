mkdir -p gpurun_out
for p in 11 10 9 8 7 6; do
  echo "=== LRF_NPROD=$p"
  for w in cfg2 distB; do
    LRF_NPROD=$p python bench.py --steps 60 --warmup 5 --workload $w --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_nprod${p}_$w.json 2>/dev/null
    python -c "
import json; d=json.load(open('gpurun_out/r2_nprod${p}_$w.json')); print('$w value %.2fM ms %.4f e2e %.2fM frame %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['frame_api'] and round(d['frame_api']['rays_per_s']/1e6,2)))"
  done
done
