#!/bin/bash
mkdir -p gpurun_out/r7
timeout 900 python -m pytest tests/test_gpu_fit.py -m gpu -q 2>&1 | tail -2
timeout 300 python tools/e2e_timing.py 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/check_multi_gpu.py 2>&1 | grep -E "identical|PASS|FAIL|rror"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r7/bench_2gpu.json 2> gpurun_out/r7/bench_2gpu.err; python -c "
import json; j=json.loads([l for l in open('gpurun_out/r7/bench_2gpu.json') if l.startswith('{')][-1]); print('2gpu ms', j['ms_per_step'], 'value', j['value'], 'e2e', j['e2e']['value'], 'share', j['roofline']['kernel_share_of_step'], 'launches', j['gpu_launches']); print(open('gpurun_out/r7/bench_2gpu.json').read()[:60])"
