#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2_gputest2.txt
tail -4 $O/r2_gputest2.txt
timeout 300 python tools/debug_gather.py > $O/r2_debug_gather2.txt 2>&1; tail -4 $O/r2_debug_gather2.txt
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_ref2.json 2> $O/r2_ref2.err
timeout 1500 python bench.py --steps 20 --warmup 5 > $O/r2_bench2.json 2> $O/r2_bench2.err
python - <<PY
import json
for f in ("r2_ref2","r2_bench2"):
    try:
        d=json.load(open("$O/%s.json"%f)); print(f, "value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], "ms/step %.3f"%d["ms_per_step"], {k:v for k,v in d.get("cpu_baseline",{}).items() if k!="sample"}, d.get("clocks"))
        if "e2e_dense_path" in d: print("  dense", d["e2e_dense_path"]); print("  e2e", {k:v for k,v in d["e2e"].items() if k!="note"}); print("  lat", d["commit_latency_ms"])
    except Exception as ex: print(f, "failed", ex)
PY
tail -5 $O/r2_bench2.err
