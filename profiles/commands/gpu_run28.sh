mkdir -p gpurun_out
( timeout 560 python -m pytest tests/ -m gpu -x -q ) > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r02_pytest_gpu_final.log | cut -c1-300
( timeout 60 python -c "import __graft_entry__ as g; g.smoke()" ) > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_smoke.log
( timeout 150 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02_bench_c2_n1_final.json 2> gpurun_out/r02_bench_c2_n1_final.err; echo "bench rc=$?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/r02_bench_c2_n1_final.json')); print(round(j['ms_per_step'],2), round(j['value']), round(j['e2e']['value']), j['config']['paf_identical_to_reference'], j['clocks'], j['roofline']['traffic'])
PY
