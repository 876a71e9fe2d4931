python -m pytest tests/test_fullsize_gpu.py -m gpu -q -x 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --no-e2e > gpurun_out/bench_r1_n2.json 2> gpurun_out/bench_r1_n2.err
python -c "import json; d=json.loads(open('gpurun_out/bench_r1_n2.json').read().strip().splitlines()[-1]); print('N=2 value=%.4e ms_per_step=%.4f kernel_ms=%.4f replay=%s'%(d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['config']['bit_exact_replay']))"; tail -3 gpurun_out/bench_r1_n2.err
