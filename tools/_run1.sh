python bench.py --steps 10 --warmup 3 > gpurun_out/bench_v8.json 2> gpurun_out/bench_v8.err
tail -c 300 gpurun_out/bench_v8.json
timeout 600 ncu --set full --clock-control none --import-source on -k regex:nasw_v3_kernel -c 15 -o gpurun_out/v3_full_v8 python bench.py --steps 1 --warmup 0 > gpurun_out/ncu_c.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 280 --csv --log-file gpurun_out/launches_v8.csv python bench.py --steps 1 --warmup 0 > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/ | tail -5
