mkdir -p gpurun_out
( timeout 500 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "nasw" ) > gpurun_out/r2_pytest_feed.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_feed.log | cut -c1-300
echo "--- one CTA of 8 warps"; timeout 120 python tools/dp_bench.py 8 100000 200 8 600 40 2>&1 | grep "^ext"
echo "--- two passes of 4 warps"; MPB_NASW_SPLIT=1 timeout 120 python tools/dp_bench.py 8 100000 200 8 600 40 2>&1 | grep "^ext"
echo "--- 350 columns: two passes of 8 warps"; timeout 120 python tools/dp_bench.py 8 100000 350 8 600 40 2>&1 | grep "^ext"
echo "--- 600 columns tb: 3 passes"; timeout 120 python tools/dp_bench.py 8 1000 24 64 20000 600 2>&1 | grep "^tb"
