timeout 600 python -m pytest tests/test_engine_gpu.py -m gpu -q -x 2>&1 | tail -3
for i in 1 2 3; do timeout 120 python bench.py --steps 30 --warmup 5 --no-e2e --no-cpu 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('  kernel_ms=%.4f frac=%.3f replay=%s'%(d['roofline']['kernel_ms'], d['roofline']['frac'], d['config']['bit_exact_replay']))"; done
