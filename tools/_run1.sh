python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for r in 1 2; do python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(round(d['value']),round(d['ms_per_step'],1),round(d['e2e']['value']),{k:round(v,1) for k,v in d['wall_ms_per_step'].items()},d['config']['paf_identical_to_reference'])"; done
python tools/dp_bench.py 592 30000 24 2>&1| head -1; python tools/dp_bench.py 16 100000 200 2>&1| head -1;  python tools/dp_bench.py 4000 10000 24 2>&1| head -2
