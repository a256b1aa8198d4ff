#!/bin/bash
# config-3 bench under different batch sizes / factor blocks per SM of the FP64 tensor-core Cholesky sweep
for b in 1024 2048 4096 16384; do for f in 2 3 4; do
  echo "batch=$b fblocks=$f"; CMFB200_DMMA_BATCH=$b CMFB200_DMMA_FBLOCKS=$f python bench.py --workload ml10m_explicit_chol_k128_f64_sideinfo --steps 3 --warmup 2 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['B_sweep_ms'], d['roofline']['A_sweep_ms'])"
done; done
