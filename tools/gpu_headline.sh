#!/bin/bash
# the two headline bench lines and the reference arm, as the driver runs them
mkdir -p gpurun_out/final2
for w in ml10m_explicit_cg_k64_f32 lastfm_implicit_cg_k64_f32; do
  timeout 900 python bench.py --workload $w > gpurun_out/final2/bench_$w.json 2> gpurun_out/final2/bench_$w.err; cut -c1-120 gpurun_out/final2/bench_$w.json
done
timeout 600 python bench.py --impl reference > gpurun_out/final2/bench_reference.json 2> gpurun_out/final2/bench_reference.err; cut -c1-200 gpurun_out/final2/bench_reference.json
