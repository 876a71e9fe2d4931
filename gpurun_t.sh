timeout 300 python -m pytest tests/test_engine_gpu.py -m gpu -q -x 2>&1 | tail -4
run() { timeout 120 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  value=%.3e kernel_ms=%.4f frac=%.3f replay=%s'%(d['value'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['bit_exact_replay']))"; }
echo bulk; run
echo nobulk; RAFTING_NO_BULK=1 run
