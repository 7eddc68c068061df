mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 2 --steps 20 --warmup 3 --workload cfg3 > gpurun_out/r2c_scale2_cfg3.json 2> gpurun_out/r2c_scale2_cfg3.err; echo "scale2 cfg3 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2c_scale2_cfg3.json')); print('N=2 cfg3 value %.2fM ms %.4f (no-exchange %.4f) e2e %.2fM launches %d exch %s' % (d['value']/1e6, d['ms_per_step'], d['config']['ms_per_step_without_exchange'], d['e2e']['value']/1e6, d['gpu_launches'], d['config']['exchange']))" || tail -8 gpurun_out/r2c_scale2_cfg3.err
