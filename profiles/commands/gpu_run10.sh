mkdir -p gpurun_out
( timeout 300 python -m pytest tests/test_gpu_stages.py -m gpu -x -q -s -k "nasw_batch_matches_oracle and pair" ) > gpurun_out/r2_pytest_pairfam.log 2>&1; echo "rc=$?"
grep -n "miniprot_b200\]\|MISMATCH\|passed\|failed" gpurun_out/r2_pytest_pairfam.log | head -20 | cut -c1-300
