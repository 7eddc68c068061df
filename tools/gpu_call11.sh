# 4 GPUs: the corrected slack protocol (weak scaling), dist test
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0,1 python -m pytest tests/test_gpu_dist.py -q -p no:cacheprovider 2>&1 | tail -3
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $1 --steps 60 --warmup 5 --scaling $2 $3 > gpurun_out/r2b_scale$1_$2$4.json 2> gpurun_out/r2b_scale$1_$2$4.err; echo "scale$1 $2 $3 rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2b_scale$1_$2$4.json"))
    print("  value %.2fM ms %.4f (no-exchange ms %s) e2e %.2fM sync %.2fM exch %s" % (d["value"]/1e6, d["ms_per_step"], d["config"]["ms_per_step_without_exchange"], d["e2e"]["value"]/1e6, d["e2e"]["sync_per_step_value"]/1e6, d["config"]["exchange"]))
except Exception as e:
    print("  no json:", e); import subprocess; print(subprocess.run(["tail","-8","gpurun_out/r2b_scale$1_$2$4.err"],capture_output=True,text=True).stdout)
PY
}
python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2b_scale1.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2b_scale1.json')); print('N=1 value %.2fM ms %.4f e2e %.2fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))"
run 4 weak "" ""
run 4 weak "--lag 2" "_lag2"
run 4 weak "--lag 0" "_lag0"
run 2 weak "" ""
