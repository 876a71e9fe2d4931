python -m pytest tests -m gpu -q -x 2>&1 | tail -5
for v in "7 2" "7 3" "6 2" "5 2"; do
  set -- $v
  touch rafting_b200/csrc/step_kernel.cuh
  RAFTING_NVCC_EXTRA="-DRAFTING_MINBLOCKS=$1 -DRAFTING_STAGES=$2" python -m rafting_b200._build > /dev/null 2>&1
  echo "VARIANT minblocks=$1 stages=$2"
  python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  value=%.3e kernel_ms=%.4f frac=%.3f replay=%s'%(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['bit_exact_replay']))"
done
touch rafting_b200/csrc/step_kernel.cuh
RAFTING_NVCC_EXTRA="-DRAFTING_MINBLOCKS=7 -DRAFTING_STAGES=2" python -m rafting_b200._build > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 12 -c 1 -o gpurun_out/prof_r1b python bench.py --steps 6 --warmup 3 --no-e2e --no-cpu > gpurun_out/prof_r1b.log 2>&1
