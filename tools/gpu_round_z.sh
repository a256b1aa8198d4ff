#!/bin/bash
S=$(date +%s)
timeout 900 python -m pytest tests/test_gpu_fit.py -m gpu -q -x 2>&1 | tail -3
timeout 600 python bench.py --workload ml10m_explicit_cg_k64_f32_implicit_features --steps 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print('cfg4 ms/step', j['ms_per_step'], 'e2e', j['e2e']['value'], 'share', j['roofline']['kernel_share_of_step'])"
echo "total $(( $(date +%s) - S )) s"
