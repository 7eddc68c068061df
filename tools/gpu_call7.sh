# 2 GPUs: exchange lag 0 vs 1, three modes; dist test again (lag plumbing)
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dist.py -q -p no:cacheprovider 2>&1 | tail -2
for mode in weak strong frame; do for lag in 1 0; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 60 --warmup 5 --scaling $mode --lag $lag > gpurun_out/r2_scale2_${mode}_lag${lag}.json 2> gpurun_out/r2_scale2_${mode}_lag${lag}.err; echo "scale2 $mode lag$lag rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_scale2_${mode}_lag${lag}.json"))
    print("  value %.2fM ms %.4f (no-exchange ms %.4f) e2e %.2fM sync %.2fM exch %s" % (d["value"]/1e6, d["ms_per_step"], d["config"]["ms_per_step_without_exchange"], d["e2e"]["value"]/1e6, d["e2e"]["sync_per_step_value"]/1e6, d["config"]["exchange"]))
except Exception as e:
    print("  no json:", e); import subprocess; print(subprocess.run(["tail","-5","gpurun_out/r2_scale2_${mode}_lag${lag}.err"],capture_output=True,text=True).stdout)
PY
done; done
python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_scale1b.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_scale1b.json')); print('N=1 value %.2fM ms %.4f e2e %.2fM sync %.2fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['sync_per_step_value']/1e6))"
