#!/bin/bash
mkdir -p gpurun_out/prof2
for w in ml10m_explicit_cg_k64_f32_implicit_features ml10m_explicit_chol_k128_f64_sideinfo ml10m_explicit_chol_k64_f32; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 400 --csv --log-file gpurun_out/prof2/launches_$w.csv \
     python bench.py --workload $w --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > /dev/null 2>&1
done
bash tools/gpu_ncu_export.sh chol_ml10m_k64 chol_sweep 2 2 -- A=1 -- python tools/quick_bench.py --shape ml10m --k 64 --solver chol --iters 1 > /dev/null 2>&1
ls gpurun_out/prof2 gpurun_out/ncu
