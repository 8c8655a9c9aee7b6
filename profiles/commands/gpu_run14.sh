mkdir -p gpurun_out
for sh in 0 1 2 3 4 5 6 7; do
  ( MPB_BENCH_SHARD=$sh timeout 120 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2_bench_shard$sh.json 2> /dev/null
  python - <<PY
import json
j=json.load(open('gpurun_out/r2_bench_shard$sh.json'))
print('shard $sh', round(j['ms_per_step'],2), {k:round(v,2) for k,v in j['wall_ms_per_step'].items()}, [(k['kernel'], round(k['ms_per_launch'],2)) for k in j['nasw_kernels'] if 'ext' in k['kernel']], j['config']['paf_identical_to_reference'])
PY
done
