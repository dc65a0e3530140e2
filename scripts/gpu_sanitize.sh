#!/bin/bash
# compute-sanitizer passes over the whole single-GPU parity suite (every kernel family, the shell_gpu_* ABI, loads, restart,
# nodal time step).  Usage (gpurun): bash scripts/gpu_sanitize.sh   -> gpurun_out/sanitize_{memcheck,racecheck,initcheck}.log
mkdir -p gpurun_out
T="tests/test_shell_gpu.py tests/test_brick_gpu.py tests/test_sh3n_gpu.py tests/test_shell_gpu_abi.py tests/test_loads_gpu.py tests/test_restart_gpu.py tests/test_dtnoda_gpu.py tests/test_ref_gpu_pin.py tests/test_forces_host_gpu.py tests/test_fail_johnson_gpu.py tests/test_balance_gpu.py"
for tool in memcheck racecheck initcheck; do
  ( timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 10 \
      python -m pytest $T -m gpu -q -k "not 1000_cycles and not large and not many_super" ; echo "rc=$?" ) > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool"; grep -E "passed|failed|SUMMARY|rc=" gpurun_out/sanitize_$tool.log | tail -4
done
