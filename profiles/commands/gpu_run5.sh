mkdir -p gpurun_out
for r in "1 30 90" "33 62 90" "70 120 60" "130 240 45" "260 480 30"; do
  ( timeout 90 python tools/pair_debug.py $r ) > "gpurun_out/r2_pairdbg_$(echo $r | tr ' ' '_').log" 2>&1
  echo "== $r rc=$?"; tail -8 "gpurun_out/r2_pairdbg_$(echo $r | tr ' ' '_').log" | cut -c1-300
done
