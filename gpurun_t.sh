timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_seglog_gpu.py -m gpu -q -x 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['e2e'], d['commit_latency_ms'])"
