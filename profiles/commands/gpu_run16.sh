mkdir -p gpurun_out
( timeout 400 python -m pytest tests/test_gpu_stages.py tests/test_gpu_configs.py -m gpu -x -q -k "nasw or c5" ) > gpurun_out/r2_pytest_mp.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2_pytest_mp.log | cut -c1-300
for sh in 7 0; do
  ( MPB_BENCH_SHARD=$sh timeout 120 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2_bench_shard$sh.json 2> gpurun_out/r2_bench_shard$sh.err || tail -3 gpurun_out/r2_bench_shard$sh.err
  python - <<PY
import json
j=json.load(open('gpurun_out/r2_bench_shard$sh.json'))
print('shard $sh', round(j['ms_per_step'],2), {k:round(v,2) for k,v in j['wall_ms_per_step'].items()}, [(k['kernel'], round(k['ms_per_launch'],2)) for k in j['nasw_kernels']], j['config']['paf_identical_to_reference'])
PY
done
