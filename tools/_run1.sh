python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print(round(d['value']),round(d['ms_per_step'],1),round(d['e2e']['value']),d['wall_ms_per_step'],d['config']['paf_identical_to_reference'])"
MPB_TRACE=1 python bench.py --steps 1 --warmup 3 2>&1 >/dev/null | grep mpb-trace | tail -19 | head -8 | cut -c1-110
python tools/dp_bench.py 592 30000 24 2>&1| head -1; python tools/dp_bench.py 16 100000 200 2>&1| head -1;  python tools/dp_bench.py 4000 10000 24 2>&1| head -2
