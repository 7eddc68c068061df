mkdir -p gpurun_out
python -m pytest tests/test_gpu_sched.py tests/test_gpu_fullsize.py -q -s -p no:cacheprovider > gpurun_out/r2_tests2.log 2>&1; echo "tests rc=$?"
grep -n "passed\|failed\|Error\|assert" gpurun_out/r2_tests2.log | tail -30
for w in cfg2 distB incoherent cfg3; do
  python bench.py --steps 60 --warmup 5 --workload $w > gpurun_out/r2_bench_${w}_a.json 2> gpurun_out/r2_bench_${w}_a.err; echo "bench $w rc=$?"; cut -c1-900 gpurun_out/r2_bench_${w}_a.json
done
python bench.py --steps 40 --warmup 5 --workload cfg5 --no-cpu-baseline > gpurun_out/r2_bench_cfg5_a.json 2> gpurun_out/r2_bench_cfg5_a.err; echo "bench cfg5 rc=$?"; cut -c1-900 gpurun_out/r2_bench_cfg5_a.json
