mkdir -p gpurun_out
( timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 ) > gpurun_out/r02_bench_c2_n2.json 2> gpurun_out/r02_bench_c2_n2.err; echo "n2 rc=$?"
tail -3 gpurun_out/r02_bench_c2_n2.err | cut -c1-300
python - <<'PY'
import json
try:
    j=json.loads([l for l in open('gpurun_out/r02_bench_c2_n2.json') if l.startswith('{')][-1]); print(j['n_gpus'], round(j['ms_per_step'],2), round(j['value']), round(j['e2e']['value']), j['config']['paf_identical_per_rank'], j['ranks']['per_rank_ms_per_step'], j['process'])
except Exception as e: print('n2 parse failed', e)
PY
( timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 ) > gpurun_out/r02_bench_c2_n2_ref.json 2> gpurun_out/r02_bench_c2_n2_ref.err; echo "n2 ref rc=$?"
tail -2 gpurun_out/r02_bench_c2_n2_ref.json | cut -c1-400
