touch rafting_b200/csrc/step_kernel.cuh
RAFTING_NVCC_EXTRA="-DRAFTING_MINBLOCKS=7 -DRAFTING_STAGES=2" python -m rafting_b200._build > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 12 -c 2 -o gpurun_out/prof_r1a python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_r1a.log 2>&1
ls -la gpurun_out/
