mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s -p no:cacheprovider > gpurun_out/r2_tests1.log 2>&1; echo "tests rc=$?" 
tail -30 gpurun_out/r2_tests1.log
python bench.py --steps 60 --warmup 5 > gpurun_out/r2_bench_cfg2_a.json 2> gpurun_out/r2_bench_cfg2_a.err; echo "bench rc=$?"; cat gpurun_out/r2_bench_cfg2_a.json | cut -c1-1500
timeout 300 python tools/umma_probe.py > gpurun_out/umma_probe.log 2>&1; echo "probe rc=$?"; tail -20 gpurun_out/umma_probe.log
