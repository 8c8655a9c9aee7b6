mkdir -p gpurun_out
# 1. launch list of the bench command (serialised, cold cache: shares only)
( timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_c2.csv python bench.py --steps 1 --warmup 3 ) > gpurun_out/r02_launches_bench.log 2>&1; echo "launch list rc=$?"; wc -l gpurun_out/r02_launches_c2.csv
# 2. full captures (no source import: the reports must stay small)
( timeout 600 ncu --set full --clock-control none -k regex:nasw_pair_kernel -c 2 -o gpurun_out/r02_pair_full -f python bench.py --steps 1 --warmup 3 ) > gpurun_out/r02_ncu_pair.log 2>&1; echo "pair rc=$?"
( timeout 600 ncu --set full --clock-control none -k regex:nasw_v3_kernel -c 8 -o gpurun_out/r02_v3_full -f python bench.py --steps 1 --warmup 3 ) > gpurun_out/r02_ncu_v3.log 2>&1; echo "v3 rc=$?"
for r in pair v3; do
  ncu -i gpurun_out/r02_${r}_full.ncu-rep --page details > gpurun_out/r02_ncu_${r}_details.txt 2>&1
  ncu -i gpurun_out/r02_${r}_full.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_alu.sum,smsp__cycles_active.avg,launch__grid_size,launch__block_size > gpurun_out/r02_ncu_${r}_raw.csv 2>&1
done
ls -la gpurun_out/*.ncu-rep
du -sm gpurun_out
for f in gpurun_out/*.ncu-rep; do s=$(stat -c %s $f); if [ $s -gt 20000000 ]; then rm -f $f; echo "removed $f ($s bytes)"; fi; done
# 3. bench lines
( MPB_TRACE=1 timeout 200 python bench.py --steps 10 --warmup 3 ) > gpurun_out/r02_bench_c2_n1.json 2> gpurun_out/r02_bench_c2_n1.err; echo "bench rc=$?"
( timeout 300 python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/r02_bench_c2_n1_ref.json 2> gpurun_out/r02_bench_c2_n1_ref.err; echo "ref rc=$?"
du -sm gpurun_out
