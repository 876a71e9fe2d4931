python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r1_n1.json 2> gpurun_out/bench_r1_n1.err
tail -c 2500 gpurun_out/bench_r1_n1.json
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_r1_ref.json 2>> gpurun_out/bench_r1_n1.err
cat gpurun_out/bench_r1_ref.json | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu > gpurun_out/launches_r1.log 2>&1
tail -3 gpurun_out/bench_r1_n1.err
