#!/bin/bash
# round-2 evidence: ncu --set full of the dominant kernel (unchanged since round 1: the capture shows it), of the new bias
# initialisation kernel and of the collective model's triangular solves; condensed CSVs only
mkdir -p gpurun_out/prof2
cap() { tag=$1; regex=$2; skip=$3; count=$4; shift 4
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c $count -o /tmp/$tag -f "$@" > gpurun_out/prof2/$tag.log 2>&1
  python tools/summarize_ncu.py /tmp/$tag.ncu-rep gpurun_out/prof2/$tag.csv; }
cap r2_ncu_full_cg_sweep_ml10m cg_resident 6 3 python tools/quick_bench.py --shape ml10m --k 64 --iters 1
cap r2_ncu_full_bias_sweep_ml10m bias_sweep 0 4 python tools/e2e_one_fit.py
cap r2_ncu_full_tri_solve_cfg4 "tri_solve|spd_factor|spmm_ones" 0 8 python bench.py --workload ml10m_explicit_cg_k64_f32_implicit_features --steps 1 --warmup 1 --no-cpu-baseline --no-e2e
ls -la gpurun_out/prof2
