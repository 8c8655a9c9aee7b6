mkdir -p gpurun_out
( MPB_BENCH_SHARD=7 MPB_TRACE=1 timeout 120 python bench.py --steps 1 --warmup 3 ) > gpurun_out/r2_bench_shard7t.json 2> gpurun_out/r2_bench_shard7t.err
grep "mpb-trace" gpurun_out/r2_bench_shard7t.err | tail -24 | cut -c1-260
