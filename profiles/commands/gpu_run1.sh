set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r2_gpu.txt; nproc >> gpurun_out/r2_gpu.txt; free -g >> gpurun_out/r2_gpu.txt
( time timeout 900 python -m pytest tests -m gpu -x -q ) > gpurun_out/r2_pytest_gpu_a.log 2>&1
( time timeout 600 python tools/parity.py C5 --json gpurun_out/r2_parity_C5.json ) > gpurun_out/r2_parity_C5.log 2>&1
( time MPB_TRACE=1 timeout 900 python tools/parity.py C4s --opt "-G 50k -e 2k" --opt "-G 50k -e 50k" --opt "-G 200k -e 2k" --opt "-G 200k -e 50k" --json gpurun_out/r2_parity_C4s.json ) > gpurun_out/r2_parity_C4s.log 2>&1
( time MPB_TRACE=1 timeout 1500 python tools/parity.py C3s --opt "-I" --json gpurun_out/r2_parity_C3s.json ) > gpurun_out/r2_parity_C3s.log 2>&1
( time timeout 600 python bench.py --steps 5 --warmup 3 ) > gpurun_out/r2_bench_a.json 2> gpurun_out/r2_bench_a.err
tail -3 gpurun_out/r2_pytest_gpu_a.log; tail -4 gpurun_out/r2_parity_C5.log;  grep identical gpurun_out/r2_parity_C4s.log | cut -c1-300; grep identical gpurun_out/r2_parity_C3s.log | cut -c1-400; tail -5 gpurun_out/r2_parity_C3s.log
