mkdir -p gpurun_out
echo "--- one CTA of 8 warps"; timeout 120 python tools/dp_bench.py 8 100000 200 8 600 40 2>&1 | grep "^ext"
echo "--- two passes of 4 warps"; MPB_NASW_SPLIT=1 timeout 120 python tools/dp_bench.py 8 100000 200 8 600 40 2>&1 | grep "^ext"
echo "--- 350 columns: two passes of 8 warps"; timeout 120 python tools/dp_bench.py 8 100000 350 8 600 40 2>&1 | grep "^ext"
echo "--- 4 warps alone, 120 columns"; timeout 120 python tools/dp_bench.py 8 100000 120 8 600 40 2>&1 | grep "^ext"
( MPB_NASW_SPLIT=1 timeout 300 ncu --set full --clock-control none --import-source on -k regex:nasw_v3_kernel -s 1 -c 1 -o gpurun_out/r02_split_mp -f python tools/dp_bench.py 8 100000 200 8 600 40 ) > gpurun_out/r02_ncu_split.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r02_split_mp.ncu-rep --page details 2>/dev/null | grep -E "Duration|Issue Slots Busy|No Eligible|Warp Cycles Per Issued|Stall|nasw_v3" | head -20
ncu -i gpurun_out/r02_split_mp.ncu-rep --page source --csv > gpurun_out/r02_split_mp_source.csv 2>/dev/null
ls -la gpurun_out/
