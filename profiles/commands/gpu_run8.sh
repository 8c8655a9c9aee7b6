mkdir -p gpurun_out
for r in "1 30 60" "33 62 60" "70 120 30" "130 240 24" "260 480 15"; do
  ( timeout 60 python tools/pair_debug.py $r 4000 ) 2>&1 | grep -E "^al|MISM" | cut -c1-200
done
for fam in pair v3; do
  for cfg in "592 30000 24 2000 3000 40" "592 30000 56 1000 20000 56" "296 30000 120 300 20000 120"; do
    ( MPB_NASW_KERNEL=$fam timeout 60 python tools/dp_bench.py $cfg ) 2>&1 | tail -2 | sed "s/^/$fam: /"
  done
done > gpurun_out/r2_dpbench_e.log 2>&1
cat gpurun_out/r2_dpbench_e.log
