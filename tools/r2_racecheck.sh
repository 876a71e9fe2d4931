#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards: the cp.async staging ring, classify_kernel's block compaction,
# the compact kernels) + synccheck + initcheck over small parity tests.  Slow tools: small cases only.
mkdir -p gpurun_out
O=gpurun_out
SEL='tests/test_engine_gpu.py::test_fuzz_parity_class_sorted_launch tests/test_engine_gpu.py::test_active_list_and_sweep_parity tests/test_compact_gpu.py::test_compact_with_unavailable_followers_three_lanes_and_escapes'
for tool in racecheck synccheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest $SEL -m gpu -x -q > $O/r2_sanitizer_$tool.log 2>&1; echo "$tool rc=$?" >> $O/r2_sanitizer_$tool.log
  echo "== $tool"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|rc=" $O/r2_sanitizer_$tool.log | tail -4
done
