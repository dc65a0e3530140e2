#!/bin/bash
# compute-sanitizer over the kernels added after profiles/r02_sanitizer.md's full run: rate-dependent three-pass loop (byte cursors,
# QEPH / BT / 3-node, table-driven copies), phase barriers, LAW36 VP = 1, Isolid 101 / 102 and co-rotational Isolid 2.
mkdir -p gpurun_out
T="tests/test_shell_gpu.py tests/test_sh3n_gpu.py tests/test_brick_gpu.py"
K="rate_dependent or vp1 or many_super_groups_rate or qeph_law36_phases or isolid_2 or formulation_variants_match_oracle"
for tool in memcheck racecheck initcheck; do
  ( timeout 400 compute-sanitizer --tool $tool --error-exitcode 9 --print-limit 10 \
      python -m pytest $T -m gpu -q -k "$K" ; echo "rc=$?" ) > gpurun_out/sanitize_new_$tool.log 2>&1
  echo "== $tool"; grep -E "passed|failed|SUMMARY|rc=" gpurun_out/sanitize_new_$tool.log | tail -4
done
