#!/bin/bash
# round-end evidence run (one GPU): parity suite, bench lines of every workload, reference arm, ncu launch lists and captures
mkdir -p gpurun_out/final
S=$(date +%s)
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -3 | tee gpurun_out/final/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | tee gpurun_out/final/smoke.txt
for w in ml10m_explicit_cg_k64_f32 lastfm_implicit_cg_k64_f32; do
  timeout 900 python bench.py --workload $w > gpurun_out/final/bench_$w.json 2> gpurun_out/final/bench_$w.err; cut -c1-160 gpurun_out/final/bench_$w.json
done
timeout 600 python bench.py --impl reference --steps 10 > gpurun_out/final/bench_reference.json 2> gpurun_out/final/bench_reference.err
for w in ml10m_explicit_cg_k128_f32 lastfm_implicit_cg_k128_f32 lastfm_implicit_cg_k256_f32 ml10m_explicit_chol_k64_f32 ml10m_explicit_cg_k64_f32_implicit_features ml10m_explicit_chol_k128_f64_sideinfo cfg1_explicit_cg_k16_f64; do
  timeout 900 python bench.py --workload $w --steps 5 --no-cpu-baseline > gpurun_out/final/bench_$w.json 2> gpurun_out/final/bench_$w.err; cut -c1-160 gpurun_out/final/bench_$w.json
done
echo "benches done $(( $(date +%s) - S )) s"
bash tools/gpu_profiles.sh > /dev/null 2>&1
ls gpurun_out/prof
echo "total $(( $(date +%s) - S )) s"
