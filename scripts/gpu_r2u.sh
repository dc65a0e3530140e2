mkdir -p gpurun_out
( timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 10 python -m pytest tests/test_forces_host_gpu.py tests/test_fail_johnson_gpu.py tests/test_balance_gpu.py -m gpu -q ; echo "rc=$?" ) > gpurun_out/sanitize_racecheck_b.log 2>&1
grep -E "passed|failed|SUMMARY|rc=" gpurun_out/sanitize_racecheck_b.log | tail -4
