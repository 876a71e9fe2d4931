python -m pytest tests -m gpu -q -x 2>&1 | tail -5
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err
tail -c 2600 gpurun_out/bench_r1_n1.json; tail -5 gpurun_out/bench_r1_n1.err
