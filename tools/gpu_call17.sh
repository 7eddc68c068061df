# round 2, last GPU call (7.7 GPU-minutes left): bf16 grid storage -- tests, bench lines (cfg2 bf16 / fp32 on the same box,
# cfg5 bf16), fp32 regression subset, one ncu capture of the bf16 kernel.  Most important first; every step bounded.
mkdir -p gpurun_out
timeout 170 python -m pytest tests/test_gpu_bf16.py -q -s -p no:cacheprovider > gpurun_out/r2_bf16_tests.log 2>&1; echo "bf16 tests rc=$?"; grep -E "passed|failed|error|bf16|Error" gpurun_out/r2_bf16_tests.log | tail -24
timeout 70 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-reference-gpu --grid-storage bf16 > gpurun_out/r2_bench_cfg2_bf16.json 2> gpurun_out/r2_bench_cfg2_bf16.err; echo "bench bf16 rc=$?"
timeout 70 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-reference-gpu > gpurun_out/r2_bench_cfg2_fp32_samebox.json 2> gpurun_out/r2_bench_cfg2_fp32.err; echo "bench fp32 rc=$?"
python - <<'PY'
import json
for n in ("cfg2_bf16", "cfg2_fp32_samebox"):
    try:
        d = json.load(open(f"gpurun_out/r2_bench_{n}.json"))
        print(n, "value %.2fM ms %.4f frac %.3f e2e %.2fM frame %.2fM clocks %s" % (d["value"]/1e6, d["ms_per_step"], d["roofline"]["frac"], d["e2e"]["value"]/1e6, (d.get("frame_api") or {}).get("rays_per_s", 0)/1e6, d["clocks"].get("sm_mhz")))
    except Exception as e:
        print(n, "unreadable", e)
PY
timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_local.py -q -p no:cacheprovider > gpurun_out/r2_fp32_subset.log 2>&1; echo "fp32 subset rc=$?"; tail -3 gpurun_out/r2_fp32_subset.log
timeout 150 python bench.py --steps 30 --warmup 5 --workload cfg5 --no-cpu-baseline --no-reference-gpu --grid-storage bf16 > gpurun_out/r2_bench_cfg5_bf16.json 2> gpurun_out/r2_bench_cfg5_bf16.err; echo "bench cfg5 bf16 rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2_bench_cfg5_bf16.json')); print('cfg5 bf16 value %.2fM ms %.4f frac %.3f' % (d['value']/1e6, d['ms_per_step'], d['roofline']['frac']))"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 6 -c 1 -f -o gpurun_out/r2_fwd_cfg2_bf16 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-gpu --grid-storage bf16 > gpurun_out/ncu_cfg2_bf16.log 2>&1; echo "ncu bf16 rc=$?"
