mkdir -p gpurun_out
( timeout 240 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -k "nasw" ) > gpurun_out/r2_pytest_pair.log 2>&1; echo "stages rc=$?"; tail -4 gpurun_out/r2_pytest_pair.log | cut -c1-300
for fam in pair v3; do
  for cfg in "592 30000 24 2000 3000 40" "592 30000 56 1000 20000 56" "296 30000 120 300 20000 120" "148 30000 240 148 20000 240"; do
    ( MPB_NASW_KERNEL=$fam timeout 60 python tools/dp_bench.py $cfg ) 2>&1 | tail -2 | sed "s/^/$fam: /"
  done
done > gpurun_out/r2_dpbench_d.log 2>&1
cat gpurun_out/r2_dpbench_d.log
( timeout 300 python -m pytest tests/test_gpu_e2e.py tests/test_gpu_dropin.py -m gpu -x -q ) > gpurun_out/r2_pytest_pair_e2e.log 2>&1; echo "e2e rc=$?"; tail -5 gpurun_out/r2_pytest_pair_e2e.log | cut -c1-300
( timeout 240 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err; echo "bench rc=$?"
cut -c1-1200 gpurun_out/r2_bench_d.json; tail -3 gpurun_out/r2_bench_d.err
