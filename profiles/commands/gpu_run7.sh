mkdir -p gpurun_out
MPB_NASW_KERNEL=pair timeout 300 ncu --set full --clock-control none --import-source on -k regex:nasw_pair_kernel -s 1 -c 1 -o gpurun_out/r2_prof_pair_ext1 python tools/dp_bench.py 592 30000 24 8 600 40 > gpurun_out/r2_ncu_pair.log 2>&1
tail -3 gpurun_out/r2_ncu_pair.log
ls -la gpurun_out/*.ncu-rep | tail -3
