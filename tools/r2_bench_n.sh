#!/bin/bash
# usage: r2_bench_n.sh N   (multi-GPU bench, both arms, as the driver launches them)
N=$1
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > $O/r2_ref_n$N.json 2> $O/r2_ref_n$N.err
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_bench_n$N.json 2> $O/r2_bench_n$N.err
python - <<PY
import json
for f in ("r2_ref_n$N","r2_bench_n$N"):
    try:
        d=json.load(open("$O/%s.json"%f)); print(f, "value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], "ms/step %.3f"%d["ms_per_step"], d["config"].get("groups_total"), d["config"].get("gather_verified"), d["scaling"])
        if "roofline" in d: print("  kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], d["clocks"], d["run"]["host_placement"], d["run"]["bit_exact_replay"]); print("  e2e", {k:v for k,v in d["e2e"].items() if k!="note"})
    except Exception as ex: print(f, "failed", ex)
PY
tail -4 $O/r2_bench_n$N.err
