#!/bin/bash
mkdir -p gpurun_out/r7
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r7/bench_${N}gpu.json 2> gpurun_out/r7/bench_${N}gpu.err; tail -3 gpurun_out/r7/bench_${N}gpu.err | cut -c1-300; cut -c1-2000 gpurun_out/r7/bench_${N}gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 10 --warmup 3 --workload lastfm_implicit_cg_k64_f32 --no-cpu-baseline > gpurun_out/r7/bench_${N}gpu_lastfm.json 2> gpurun_out/r7/bench_${N}gpu_lastfm.err; tail -3 gpurun_out/r7/bench_${N}gpu_lastfm.err | cut -c1-300; cut -c1-400 gpurun_out/r7/bench_${N}gpu_lastfm.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --impl reference --gpus $N --steps 2 --warmup 1 | cut -c1-200
