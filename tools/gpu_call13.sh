mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_dist.py -q -p no:cacheprovider 2>&1 | tail -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 60 --warmup 5 > gpurun_out/r2c_scale2_default.json 2> gpurun_out/r2c_scale2_default.err; echo "scale2 default rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2c_scale2_default.json')); print('N=2 value %.2fM ms %.4f (no-exchange %.4f) e2e %.2fM exch %s' % (d['value']/1e6, d['ms_per_step'], d['config']['ms_per_step_without_exchange'], d['e2e']['value']/1e6, d['config']['exchange']))" || tail -5 gpurun_out/r2c_scale2_default.err
