#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > $O/r2_gputest3.txt
tail -4 $O/r2_gputest3.txt
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_ref3.json 2> $O/r2_ref3.err
timeout 600 python tools/bench_secondary.py > $O/r2_secondary3.jsonl 2> $O/r2_secondary3.err
python - <<PY
import json
try:
    d=json.load(open("$O/r2_ref3.json")); print("ref3 value %.4g"%d["value"], {k:v for k,v in d["cpu_baseline"].items() if k!="sample"})
except Exception as ex: print("ref failed", ex)
for l in open("$O/r2_secondary3.jsonl"):
    try:
        r=json.loads(l); print(r["config"], r["kernel_ms_per_step_median"], r["roofline"]["frac"], r["rates_per_s"])
    except Exception as ex: print("sec parse", ex)
PY
tail -3 $O/r2_secondary3.err
