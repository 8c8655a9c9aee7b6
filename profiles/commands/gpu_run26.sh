mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "nasw_batch or long_wide" ) > gpurun_out/r2_pytest_feed2.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_feed2.log | cut -c1-300
echo "--- two passes of 4 warps"; MPB_NASW_SPLIT=1 timeout 120 python tools/dp_bench.py 8 100000 200 8 600 40 2>&1 | grep "^ext"
echo "--- 350 columns: two passes of 8 warps"; timeout 120 python tools/dp_bench.py 8 100000 350 8 600 40 2>&1 | grep "^ext"
echo "--- 350 columns: three passes of 4 warps"; MPB_NASW_WIDE_WARPS=4 timeout 120 python tools/dp_bench.py 8 100000 350 8 600 40 2>&1 | grep "^ext"
run() { tag=$1; shift; ( env "$@" MPB_TRACE=1 timeout 150 python bench.py --steps 6 --warmup 3 ) > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err; python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2_bench_$tag.json')); print('$tag', round(j['ms_per_step'],2), round(j['wall_ms_per_step']['S3_dp_waves'],2), j['config']['paf_identical_to_reference'])
except Exception as e: print('$tag failed', e)
PY
}
run s7 MPB_BENCH_SHARD=7
run s7w4 MPB_BENCH_SHARD=7 MPB_NASW_WIDE_WARPS=4
run s0split MPB_NASW_SPLIT=1
