#!/bin/bash
# hot opposing rows in shared memory (CMFB200_RES_HOT): parity of the CG sweeps, then ms per iteration with / without
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_sweeps.py tests/test_gpu_bench_shapes.py -q -m gpu -x -k "float32 or full_shape or large_rank or f32" 2>&1 | tail -5 > gpurun_out/r2s3_hot_t1.log
cat gpurun_out/r2s3_hot_t1.log
for w in ml10m_explicit_cg_k64_f32 lastfm_implicit_cg_k64_f32 ml10m_explicit_cg_k128_f32; do
  for hot in 0 1; do
    CMFB200_RES_HOT=$hot timeout 300 python bench.py --workload $w --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); r=d['roofline']; print('$w hot=$hot ms %.3f B %.3f A %.3f' % (d['ms_per_step'], r.get('B_sweep_ms',0), r.get('A_sweep_ms',0)))"
  done
done 2>&1 | tee gpurun_out/r2s3_hot_bench.log
