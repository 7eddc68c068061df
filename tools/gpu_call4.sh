# 2 GPUs: sharded-render test, scaling bench at N=2 in the three modes and both exchanges
mkdir -p gpurun_out
python -m pytest tests/test_gpu_dist.py tests/test_gpu_pipeline.py tests/test_gpu_renderer_loop.py -q -s -p no:cacheprovider > gpurun_out/r2_tests4.log 2>&1; echo "tests rc=$?"; tail -25 gpurun_out/r2_tests4.log
for mode in weak strong frame; do for ex in fused nccl; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 40 --warmup 5 --scaling $mode --exchange $ex > gpurun_out/r2_scale2_${mode}_${ex}.json 2> gpurun_out/r2_scale2_${mode}_${ex}.err; echo "scale2 $mode $ex rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_scale2_${mode}_${ex}.json"))
    print("  value %.2fM ms %.3f e2e %.2fM sync %.2fM exch %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["e2e"]["sync_per_step_value"]/1e6, d["config"]["exchange"]))
except Exception as e:
    print("  no json:", e); import subprocess; print(subprocess.run(["tail","-5","gpurun_out/r2_scale2_${mode}_${ex}.err"],capture_output=True,text=True).stdout)
PY
done; done
python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_scale1.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_scale1.json')); print('N=1 value %.2fM e2e %.2fM sync %.2fM' % (d['value']/1e6, d['e2e']['value']/1e6, d['e2e']['sync_per_step_value']/1e6))"
