mkdir -p gpurun_out
python -m pytest tests/test_gpu_sched.py tests/test_gpu_pipeline.py -q -p no:cacheprovider > gpurun_out/r2_tests3.log 2>&1; echo "tests rc=$?"; tail -5 gpurun_out/r2_tests3.log
python tools/host_overhead.py > gpurun_out/host_overhead.log 2>&1; head -40 gpurun_out/host_overhead.log
echo "=== TC backward"
LRF_BWD_TC=1 timeout 600 python -m pytest tests/test_gpu_backward.py tests/test_gpu_grads.py -q -x -p no:cacheprovider > gpurun_out/r2_tc_bwd_tests.log 2>&1; echo "tc tests rc=$?"; tail -30 gpurun_out/r2_tc_bwd_tests.log
echo "=== train time (CUDA-core shade)"
timeout 300 python tools/train_time.py 2>&1 | tail -4
echo "=== train time (tcgen05 shade)"
LRF_BWD_TC=1 timeout 300 python tools/train_time.py 2>&1 | tail -4
