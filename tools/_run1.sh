python -m pytest tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -2
for r in 1 2; do python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(round(d['value']),round(d['ms_per_step'],1),round(d['e2e']['value']),{k:round(v,1) for k,v in d['wall_ms_per_step'].items()},d['config']['paf_identical_to_reference'])"; done
