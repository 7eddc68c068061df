mkdir -p gpurun_out
timeout 240 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/r2_bench_final.json
timeout 200 python bench.py --workload cfg5 --steps 40 --no-cpu-baseline > gpurun_out/r2_bench_cfg5_final.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_cfg5_final.json')); print('cfg5 value %.2fM ms %.4f e2e %.2fM frac %.3f' % (d['value']/1e6, d['ms_per_step'], d['e2e']['value']/1e6, d['roofline']['frac']))"
timeout 300 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_backward.py > gpurun_out/r2_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|worst gradient|schedule kernels|pipeline frames" gpurun_out/r2_sanitizer_memcheck.txt
timeout 200 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_small.py > gpurun_out/r2_sanitizer_fwd_memcheck.txt 2>&1; echo "memcheck fwd rc=$?"; grep -E "ERROR SUMMARY|rgb err" gpurun_out/r2_sanitizer_fwd_memcheck.txt
