mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -q -x -p no:cacheprovider > gpurun_out/r2_pe_tests.log 2>&1; echo "parity(+PE) rc=$?"; tail -15 gpurun_out/r2_pe_tests.log
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider --deselect tests/test_gpu_parity.py > gpurun_out/r2_final_tests.log 2>&1; echo "all gpu tests rc=$?"; tail -8 gpurun_out/r2_final_tests.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
