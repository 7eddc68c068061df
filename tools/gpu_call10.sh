# 8 GPUs: the driver's scaling command (weak, fused exchange, lag 1) at N = 8, 4, 1 on the same box + strong / frame at 8
mkdir -p gpurun_out
run() {  # n mode extra
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $1 --steps 60 --warmup 5 --scaling $2 $3 > gpurun_out/r2_scale$1_$2$4.json 2> gpurun_out/r2_scale$1_$2$4.err; echo "scale$1 $2 $3 rc=$?"
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2_scale$1_$2$4.json"))
    print("  value %.2fM ms %.4f (no-exchange ms %s) e2e %.2fM sync %.2fM exch %s clocks %s" % (d["value"]/1e6, d["ms_per_step"], d["config"]["ms_per_step_without_exchange"], d["e2e"]["value"]/1e6, d["e2e"]["sync_per_step_value"]/1e6, d["config"]["exchange"], d["clocks"].get("sm_mhz")))
except Exception as e:
    print("  no json:", e); import subprocess; print(subprocess.run(["tail","-8","gpurun_out/r2_scale$1_$2$4.err"],capture_output=True,text=True).stdout)
PY
}
python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_scale1_box8.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_scale1_box8.json')); print('N=1 value %.2fM ms %.4f e2e %.2fM' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6))"
run 8 weak "" ""
run 8 weak "--exchange nccl" "_nccl"
run 4 weak "" ""
run 2 weak "" ""
run 8 strong "" ""
run 8 frame "" ""
