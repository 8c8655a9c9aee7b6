mkdir -p gpurun_out
for fam in auto v3 pair auto v3; do
  ( MPB_NASW_KERNEL=$fam timeout 200 python bench.py --steps 8 --warmup 3 ) > gpurun_out/r2_bench_ab_$fam.json 2> /dev/null
  python - <<PY
import json
j=json.load(open('gpurun_out/r2_bench_ab_$fam.json'))
print('$fam', round(j['ms_per_step'],2), {k:round(v,2) for k,v in j['wall_ms_per_step'].items()}, [(k['kernel'], round(k['ms_per_launch'],2)) for k in j['nasw_kernels']])
PY
done
