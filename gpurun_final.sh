timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1.err
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_n1.json 2>> gpurun_out/bench_r1.err
python -c "
import json
for f in ('gpurun_out/bench_r1_ref.json','gpurun_out/bench_r1_n1.json'):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, '%.4e'%d['value'], d.get('e2e',{}).get('value'), d.get('roofline',{}).get('frac'), d.get('cpu_baseline',{}).get('value'), d.get('clocks'))
"
tail -3 gpurun_out/bench_r1.err
