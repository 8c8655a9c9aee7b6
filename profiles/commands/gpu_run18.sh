mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_stages.py tests/test_gpu_e2e.py -m gpu -x -q -k "sort or seed or refine or chain or tiny or small" ) > gpurun_out/r2_pytest_sort.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_sort.log | cut -c1-300
( MPB_TRACE=1 timeout 400 python tools/parity.py C3s --opt=-I --opt=-I --json gpurun_out/r2_parity_C3s_b.json ) > gpurun_out/r2_parity_C3s_b.log 2>&1; echo "C3s rc=$?"
python - <<'PY'
import json
for r in json.load(open('gpurun_out/r2_parity_C3s_b.json')): print({k:r[k] for k in ('identical','ours_s','ref_map_s','wall_ms','anchors_per_protein')})
PY
( timeout 120 python bench.py --steps 6 --warmup 3 ) > gpurun_out/r2_bench_g.json 2> gpurun_out/r2_bench_g.err || tail -3 gpurun_out/r2_bench_g.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_bench_g.json'))
print(round(j['ms_per_step'],2), j['stage_ms_per_step'], {k:round(v,2) for k,v in j['wall_ms_per_step'].items()}, j['config']['paf_identical_to_reference'])
PY
