mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2_final_tests2.log 2>&1; echo "all gpu tests rc=$?"; tail -4 gpurun_out/r2_final_tests2.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_final20.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_final20.json')); print('value %.2fM ms %.4f e2e %.2fM sync %.2fM refgpu %s' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['e2e']['sync_per_step_value']/1e6, d['reference_gpu']))"
