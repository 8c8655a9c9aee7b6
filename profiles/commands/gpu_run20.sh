mkdir -p gpurun_out
# 0. index build on the device, splice scores
( timeout 500 python -m pytest tests/test_gpu_dropin.py -m gpu -x -q ) > gpurun_out/r2_pytest_idx.log 2>&1; rc=$?; echo "pytest dropin rc=$rc"; tail -12 gpurun_out/r2_pytest_idx.log | cut -c1-300
if [ $rc -ne 0 ]; then export MPB_IDX_BUILD=host; echo "falling back to host index build for the measurements"; fi
# 1. launch list of the bench command (serialised, cold cache: shares only)
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 1 --warmup 3 ) > gpurun_out/r02_launches_bench.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/r02_launches_c2.csv
# 2. full capture of the traceback pair kernel and of the extension kernels, first step of the same command
( timeout 900 ncu --set full --clock-control none --import-source on -k regex:'nasw_pair_kernel|nasw_v3_kernel' -c 16 -o gpurun_out/r02_nasw_full -f python bench.py --steps 1 --warmup 3 ) > gpurun_out/r02_ncu_full.log 2>&1; echo "full rc=$?"; ls -la gpurun_out/r02_nasw_full.ncu-rep
# 3. a clean bench line with the wave trace
( MPB_TRACE=1 timeout 200 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02_bench_c2_n1.json 2> gpurun_out/r02_bench_c2_n1.err; echo "bench rc=$?"
( timeout 300 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r02_bench_c2_n1_ref.json 2> gpurun_out/r02_bench_c2_n1_ref.err; echo "ref rc=$?"
python - <<'PY'
import json
j=json.load(open('gpurun_out/r02_bench_c2_n1.json')); print(round(j['ms_per_step'],2), j['e2e']['value'], j['clocks'])
j=json.load(open('gpurun_out/r02_bench_c2_n1_ref.json')); print(j['value'], j['ms_per_step'])
PY
