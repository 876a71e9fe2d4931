ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 12 -c 1 -o gpurun_out/prof_r1g python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_r1g.log 2>&1
