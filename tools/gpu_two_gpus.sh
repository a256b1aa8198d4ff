#!/bin/bash
mkdir -p gpurun_out/r6
S=$(date +%s)
timeout 300 python tools/e2e_timing.py 2>&1 | tail -9
timeout 600 python -m pytest tests/test_gpu_fit.py -m gpu -q -k "initial_state or explicit_fit" 2>&1 | tail -2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi_gpu.py 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r6/bench_2gpu.json 2> gpurun_out/r6/bench_2gpu.err; tail -3 gpurun_out/r6/bench_2gpu.err; cut -c1-1500 gpurun_out/r6/bench_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --workload lastfm_implicit_cg_k64_f32 --no-cpu-baseline > gpurun_out/r6/bench_2gpu_lastfm.json 2> gpurun_out/r6/bench_2gpu_lastfm.err; tail -3 gpurun_out/r6/bench_2gpu_lastfm.err; cut -c1-600 gpurun_out/r6/bench_2gpu_lastfm.json
echo "total $(( $(date +%s) - S )) s"
