mkdir -p gpurun_out
echo "=== TC backward (hi/lo wgrad)"
LRF_BWD_TC=1 timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_grads.py -q -p no:cacheprovider > gpurun_out/r2_tc_bwd_tests2.log 2>&1; echo "tc tests rc=$?"; tail -15 gpurun_out/r2_tc_bwd_tests2.log
LRF_BWD_TC=1 timeout 300 python tools/train_time.py 2>&1 | tail -4
echo "=== ncu forward cfg2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 6 -c 1 -f -o gpurun_out/r2_fwd_cfg2 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_cfg2.log 2>&1; echo "ncu cfg2 rc=$?"
echo "=== ncu forward cfg5 (640^3)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_kernel -s 8 -c 1 -f -o gpurun_out/r2_fwd_cfg5 python bench.py --steps 3 --warmup 3 --workload cfg5 --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_cfg5.log 2>&1; echo "ncu cfg5 rc=$?"
echo "=== launch list cfg2"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_cfg2.csv python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/ncu_list.log 2>&1; echo "list rc=$?"
echo "=== ncu TC shade"
LRF_BWD_TC=1 ITERS=2 timeout 600 ncu --set full --clock-control none --import-source on -k regex:bwd_shade_tc -s 1 -c 1 -f -o gpurun_out/r2_bwd_shade_tc python tools/bwd_profile.py > gpurun_out/ncu_bwd.log 2>&1; echo "ncu bwd rc=$?"
LRF_BWD_TC=1 ITERS=3 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_launches_bwd_tc.csv python tools/bwd_profile.py > /dev/null 2>&1; echo "bwd list rc=$?"
ls -la gpurun_out/*.ncu-rep
