#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader > $O/r2_n2_gpus.txt; nvidia-smi topo -m >> $O/r2_n2_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_seglog_gpu.py tests/test_compact_gpu.py -m gpu -x -q 2>&1 | tail -15 > $O/r2_gputest_n2.txt
tail -4 $O/r2_gputest_n2.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2_bench_n2.json 2> $O/r2_bench_n2.err
python - <<PY
import json
try:
    d=json.load(open("$O/r2_bench_n2.json")); print("N=2 value %.4g"%d["value"], "e2e %.4g"%d["e2e"]["value"], "ms/step %.3f"%d["ms_per_step"], d["config"], {k:v for k,v in d["run"].items() if k!="inputs"}, d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["clocks"])
    print({k:v for k,v in d["e2e"].items() if k!="note"})
except Exception as ex: print("bench n2 failed", ex)
PY
tail -5 $O/r2_bench_n2.err
