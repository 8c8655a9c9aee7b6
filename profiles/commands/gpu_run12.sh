mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_g.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_gpu_g.log | cut -c1-300
for n in 74 148 296; do ( MPB_NASW_KERNEL=pair timeout 60 python tools/dp_bench.py $n 30000 120 8 600 40 ) 2>&1 | grep "^ext" | sed "s/^/pair n=$n: /"; done
for n in 74 148 296; do ( MPB_NASW_KERNEL=v3 timeout 60 python tools/dp_bench.py $n 30000 120 8 600 40 ) 2>&1 | grep "^ext" | sed "s/^/v3 n=$n: /"; done
