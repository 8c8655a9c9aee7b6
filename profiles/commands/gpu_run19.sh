mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_stages.py tests/test_gpu_e2e.py -m gpu -x -q ) > gpurun_out/r2_pytest_spsc.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2_pytest_spsc.log | cut -c1-400
( timeout 120 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2_bench_h.json 2> gpurun_out/r2_bench_h.err || tail -3 gpurun_out/r2_bench_h.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_bench_h.json'))
print(round(j['ms_per_step'],2), j['config']['paf_identical_to_reference'])
PY
