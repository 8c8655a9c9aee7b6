set -x
mkdir -p gpurun_out
( time MPB_TRACE=1 timeout 1500 python tools/parity.py C3s --opt=-I --keep --json gpurun_out/r2_parity_C3s.json ) > gpurun_out/r2_parity_C3s.log 2>&1
( time MPB_TRACE=1 timeout 300 python tools/parity.py C2 --opt= --opt= --json gpurun_out/r2_parity_C2.json ) > gpurun_out/r2_parity_C2.log 2>&1
grep identical gpurun_out/r2_parity_C3s.log | cut -c1-600; tail -5 gpurun_out/r2_parity_C3s.log
