set -x
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_configs.py tests/test_gpu_stages.py -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_b.log 2>&1
( time timeout 600 python bench.py --impl reference --steps 5 --warmup 2 ) > gpurun_out/r2_bench_ref_b.json 2> gpurun_out/r2_bench_ref_b.err
( time timeout 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2_bench_b.json 2> gpurun_out/r2_bench_b.err
tail -5 gpurun_out/r2_pytest_gpu_b.log; cat gpurun_out/r2_bench_ref_b.json | cut -c1-300; tail -3 gpurun_out/r2_bench_b.err; cat gpurun_out/r2_bench_b.json | cut -c1-600
