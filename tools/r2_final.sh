#!/bin/bash
# what the driver runs at round end, in one call: smoke, the GPU tests, both bench arms at N = 1
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.txt 2>&1; tail -2 $O/r2_smoke.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > $O/r2_gputest_final.txt; tail -3 $O/r2_gputest_final.txt
timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $O/r2_ref_final.json 2> $O/r2_ref_final.err
timeout 1500 python bench.py --gpus 1 --steps 20 --warmup 5 > $O/r2_bench_final.json 2> $O/r2_bench_final.err
python - <<PY
import json
for f in ("r2_ref_final","r2_bench_final"):
    try:
        d=json.load(open("$O/%s.json"%f)); print(f, "value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], "ms/step %.3f"%d["ms_per_step"], {k:v for k,v in d.get("cpu_baseline",{}).items() if k!="sample"})
        if "roofline" in d: print("  kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], d["clocks"]); print("  e2e", {k:v for k,v in d["e2e"].items() if k!="note"}); print("  frames", d.get("e2e_from_frames")); print("  dense", d.get("e2e_dense_path")); print("  lat", d["commit_latency_ms"]); print("  sec", [(s["config"], s["kernel_ms_per_step"], s["roofline_frac"]) for s in d.get("secondary_rates",[])])
    except Exception as ex: print(f, "failed", ex)
PY
tail -3 $O/r2_bench_final.err
