#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_compact_gpu.py tests/test_engine_gpu.py -m gpu -x -q 2>&1 | tail -6 > $O/r2_gputest5.txt; tail -3 $O/r2_gputest5.txt
timeout 1500 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary > $O/r2_bench5.json 2> $O/r2_bench5.err
RAFTING_BULK_STAGING=1 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary --no-e2e > $O/r2_bench5_tma.json 2> $O/r2_bench5_tma.err
python - <<PY
import json
for f in ("r2_bench5","r2_bench5_tma"):
    try:
        d=json.load(open("$O/%s.json"%f)); print(f, "value %.4g"%d["value"], "kernel_ms %.5f"%d["roofline"]["kernel_ms"], "frac %.4f"%d["roofline"]["frac"], d["run"]["bit_exact_replay"])
        if "e2e" in d: print("  e2e", {k:v for k,v in d["e2e"].items() if k!="note"}); print("  dense", d.get("e2e_dense_path")); print("  lat", d["commit_latency_ms"])
    except Exception as ex: print(f, "failed", ex)
PY
tail -3 $O/r2_bench5.err
# ncu summaries of the kernels round 2 added (one GPU, short commands)
cat > /tmp/run_cfg.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tools"))
import bench_secondary as b
L = b._bind()
if sys.argv[1] == "5": b.run("config5", 524288, 3, 4, 6, 0x5EED0005, L.rafting_wl_mixed_step, elect=True, pool=True, quiet=True)
else: b.run("config3", 262144, 5, 1, 6, 0x5EED0003, L.rafting_wl_vote_step, local_slot=2, quiet=True)
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slow_kernel -s 6 -c 1 -f -o $O/prof_r2b_slow_cfg5 python /tmp/run_cfg.py 5 > $O/prof_r2b_slow_cfg5.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:slow_kernel -s 6 -c 1 -f -o $O/prof_r2b_slow_cfg3 python /tmp/run_cfg.py 3 > $O/prof_r2b_slow_cfg3.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:classify_kernel -s 6 -c 1 -f -o $O/prof_r2b_classify python /tmp/run_cfg.py 5 > $O/prof_r2b_classify.log 2>&1
timeout 900 ncu --set full --clock-control none -k "regex:unpack_kernel|pack_kernel" -s 40 -c 2 -f -o $O/prof_r2b_compact python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary --launches 8 > $O/prof_r2b_compact.log 2>&1
ls -la $O/prof_r2b_*.ncu-rep
