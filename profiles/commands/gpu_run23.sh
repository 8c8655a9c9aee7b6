mkdir -p gpurun_out
( timeout 500 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "two_passes" ) > gpurun_out/r2_sanitizer_split.log 2>&1; echo "rc=$?"
grep -A14 "Invalid\|misaligned\|Misaligned" gpurun_out/r2_sanitizer_split.log | head -50
