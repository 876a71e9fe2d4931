#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
for g in 524288 1048576; do
timeout 600 ncu --set full --clock-control none -k regex:step_kernel -s 25 -c 1 -f -o $O/prof_r2c_g$g python bench.py --groups $g --steps 1 --warmup 3 --no-cpu --no-e2e --no-secondary --launches 4 > $O/prof_r2c_g$g.log 2>&1
timeout 300 python bench.py --groups $g --steps 10 --warmup 3 --no-cpu --no-e2e --no-secondary --launches 8 > $O/r2_bench_g$g.json 2> $O/r2_bench_g$g.err
done
ls -la $O/prof_r2c_*.ncu-rep
python - <<PY
import json
for g in (524288, 1048576):
    try:
        d=json.load(open("$O/r2_bench_g%d.json"%g)); print(g, "value %.4g"%d["value"], "kernel_ms %.4f"%d["roofline"]["kernel_ms"], "frac %.3f"%d["roofline"]["frac"])
    except Exception as ex: print(g, "failed", ex)
PY
