#!/bin/bash
# on the GPU box: gather debug, then one ncu --set full capture of each hot-kernel variant (one GPU, short command)
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python tools/debug_gather.py 262144 > $O/r2_debug_gather.txt 2>&1; tail -4 $O/r2_debug_gather.txt
CMD="python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-secondary --launches 6"
RAFTING_NO_PAIR=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 30 -c 2 -f -o $O/prof_r2a_v6 $CMD > $O/prof_r2a_v6.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pair_kernel -s 30 -c 2 -f -o $O/prof_r2a_pair $CMD > $O/prof_r2a_pair.log 2>&1
ls -la $O/prof_r2a_*.ncu-rep
