#!/bin/bash
# two GPUs: the 2-GPU fit against the 1-GPU fit (pytest, tools/check_multi_gpu.py under torchrun), then the default workload
mkdir -p gpurun_out/r2final
O=gpurun_out/r2final
timeout 600 python -m pytest tests/test_gpu_bench_shapes.py -q -m gpu -k two_gpu 2>&1 | tail -3 > $O/two_gpu_test.log; cat $O/two_gpu_test.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/check_multi_gpu.py > $O/multi_gpu_check_2gpu.log 2>&1; grep -E "identical|PASS|FAIL|rror" $O/multi_gpu_check_2gpu.log | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > $O/bench_2gpu.json 2> $O/bench_2gpu.err
python -c "
import json; j=json.loads([l for l in open('$O/bench_2gpu.json') if l.startswith('{')][-1]); print('2gpu ms', j['ms_per_step'], 'value', j['value'], 'e2e', j['e2e']['value'], 'share', j['roofline']['kernel_share_of_step'])"
