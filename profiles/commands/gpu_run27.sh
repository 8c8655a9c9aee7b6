mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "nasw" ) > gpurun_out/r2_pytest_w4.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_w4.log | cut -c1-300
echo "--- 600 columns tb"; timeout 120 python tools/dp_bench.py 8 1000 24 64 20000 600 2>&1 | grep "^tb"
echo "--- 600 columns tb, 8-warp passes"; MPB_NASW_WIDE_WARPS=8 timeout 120 python tools/dp_bench.py 8 1000 24 64 20000 600 2>&1 | grep "^tb"
( timeout 300 python tools/parity.py C5 --json gpurun_out/r2_parity_C5_b.json ) > gpurun_out/r2_parity_C5_b.log 2>&1; echo "C5 rc=$?"
python - <<'PY'
import json
for r in json.load(open('gpurun_out/r2_parity_C5_b.json')): print({k:r[k] for k in ('identical','ours_s','ref_map_s','wall_ms')})
PY
( timeout 150 python bench.py --steps 8 --warmup 3 ) > gpurun_out/r2_bench_w4.json 2> gpurun_out/r2_bench_w4.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_bench_w4.json')); print('default', round(j['ms_per_step'],2), j['config']['paf_identical_to_reference'])
PY
