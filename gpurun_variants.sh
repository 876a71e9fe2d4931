run() { python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  value=%.3e kernel_ms=%.4f frac=%.3f replay=%s'%(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['bit_exact_replay']))"; }
for v in "64 8" "32 16" "128 4" "64 7" "96 5"; do
  set -- $v
  touch rafting_b200/csrc/step_kernel.cuh
  RAFTING_NVCC_EXTRA="-DRAFTING_TPB=$1 -DRAFTING_MINBLOCKS=$2" python -m rafting_b200._build > /dev/null 2>&1
  echo "VARIANT tpb=$1 minblocks=$2"; run
done
