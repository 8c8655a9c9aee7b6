mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "nasw" ) > gpurun_out/r2_pytest_split.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_split.log | cut -c1-300
run() { tag=$1; shift; ( env "$@" MPB_TRACE=1 timeout 150 python bench.py --steps 8 --warmup 3 ) > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err; python - <<PY
import json
try:
    j=json.load(open('gpurun_out/r2_bench_$tag.json')); print('$tag', round(j['ms_per_step'],2), round(j['wall_ms_per_step']['S3_dp_waves'],2), j['config']['paf_identical_to_reference'])
except Exception as e: print('$tag failed', e)
PY
grep "S3: wave1" gpurun_out/r2_bench_$tag.err | tail -2; }
run split X=1
run nosplit MPB_NASW_SPLIT=0
run split_pair MPB_NASW_KERNEL=pair
grep "mpb-trace\] nasw ext class\|tb  class" gpurun_out/r2_bench_split_pair.err | tail -9 | cut -c1-120
grep "mpb-trace\] nasw ext class" gpurun_out/r2_bench_split.err | tail -8 | cut -c1-120
