set -x
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "nasw" ) > gpurun_out/r2_pytest_pair.log 2>&1
tail -15 gpurun_out/r2_pytest_pair.log
for fam in pair v3; do
  ( MPB_NASW_KERNEL=$fam timeout 300 python tools/dp_bench.py 592 30000 24 2000 3000 40 ) > gpurun_out/r2_dpbench_${fam}_24.log 2>&1
  ( MPB_NASW_KERNEL=$fam timeout 300 python tools/dp_bench.py 592 30000 56 1000 20000 56 ) > gpurun_out/r2_dpbench_${fam}_56.log 2>&1
  ( MPB_NASW_KERNEL=$fam timeout 300 python tools/dp_bench.py 296 30000 120 300 20000 120 ) > gpurun_out/r2_dpbench_${fam}_120.log 2>&1
  ( MPB_NASW_KERNEL=$fam timeout 300 python tools/dp_bench.py 148 30000 240 148 20000 240 ) > gpurun_out/r2_dpbench_${fam}_240.log 2>&1
  tail -n 2 gpurun_out/r2_dpbench_${fam}_*.log
done
( time timeout 600 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_dropin.py -m gpu -x -q ) > gpurun_out/r2_pytest_pair_e2e.log 2>&1
tail -5 gpurun_out/r2_pytest_pair_e2e.log
( time timeout 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2_bench_c.json 2> gpurun_out/r2_bench_c.err
cut -c1-1500 gpurun_out/r2_bench_c.json; tail -3 gpurun_out/r2_bench_c.err
