mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_f.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest_gpu_f.log | cut -c1-300
( timeout 240 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2_bench_f.json 2> gpurun_out/r2_bench_f.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_f.err | cut -c1-300
( timeout 240 python bench.py --impl reference --steps 5 --warmup 2 ) > gpurun_out/r2_bench_ref_f.json 2> gpurun_out/r2_bench_ref_f.err; echo "ref rc=$?"
