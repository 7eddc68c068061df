# last seconds of the round's GPU budget: memcheck of the bf16 instantiation on a tiny case
mkdir -p gpurun_out
timeout 60 compute-sanitizer --tool memcheck --print-limit 5 python tools/sanitize_bf16.py > gpurun_out/r2_sanitizer_bf16.txt 2>&1; echo "memcheck bf16 rc=$?"; grep -E "ERROR SUMMARY|bf16 ==" gpurun_out/r2_sanitizer_bf16.txt
