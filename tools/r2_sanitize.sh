#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_compact_gpu.py tests/test_seglog_gpu.py "tests/test_engine_gpu.py::test_shard_image_survives_the_process" tests/test_multigpu_gpu.py -m gpu -x -q -k "not config4" > $O/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> $O/r2_sanitizer_memcheck.log
tail -12 $O/r2_sanitizer_memcheck.log
cat > /tmp/run_cfg.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import bench_secondary as b
L = b._bind()
b.run("config5", 32768, 3, 4, 8, 0x5EED0005, L.rafting_wl_mixed_step, elect=True, pool=True, quiet=True)
b.run("config3", 16384, 5, 1, 6, 0x5EED0003, L.rafting_wl_vote_step, local_slot=2, quiet=True)
print("streams ok")
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/run_cfg.py > $O/r2_sanitizer_streams.log 2>&1; echo "memcheck rc=$?" >> $O/r2_sanitizer_streams.log
tail -5 $O/r2_sanitizer_streams.log
