#!/bin/bash
# on the GPU box: parity first (bounded by timeouts), then the A/B of the hot kernel, then both bench arms
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/r2_gputest.txt
tail -3 $O/r2_gputest.txt
for v in pair v6; do
  if [ $v = v6 ]; then export RAFTING_NO_PAIR=1; else unset RAFTING_NO_PAIR; fi
  timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu --no-secondary --no-e2e > $O/r2_ab_$v.json 2> $O/r2_ab_$v.err
  python - <<PY
import json
try:
    d=json.load(open("$O/r2_ab_$v.json")); print("$v", "value %.3g"%d["value"], "kernel_ms %.4f"%d["roofline"]["kernel_ms"], "pair_ms %.4f"%d["roofline"]["kernel_ms_event_pair_per_launch"], "frac %.3f"%d["roofline"]["frac"], "replay", d["run"]["bit_exact_replay"], "region_ms %.1f"%d["timed_region_ms"])
except Exception as ex: print("$v failed", ex)
PY
done
unset RAFTING_NO_PAIR
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_ref.json 2> $O/r2_ref.err
timeout 1200 python bench.py --steps 20 --warmup 5 > $O/r2_bench.json 2> $O/r2_bench.err
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/r2_ref_b.json 2>> $O/r2_ref.err
python - <<PY
import json
for f in ("r2_ref","r2_bench","r2_ref_b"):
    try:
        d=json.load(open("$O/%s.json"%f)); print(f, "value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], "ms/step %.3f"%d["ms_per_step"], d.get("cpu_baseline",{}).get("value"), d.get("clocks"))
    except Exception as ex: print(f, "failed", ex)
PY
tail -3 $O/r2_bench.err
