#!/bin/bash
# short round-end verification: parity suite, smoke(), the two headline bench lines
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
for w in ml10m_explicit_cg_k64_f32 lastfm_implicit_cg_k64_f32; do
  timeout 900 python bench.py --workload $w 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print(j['config']['workload'], 'ms', round(j['ms_per_step'],3), 'value', '%.3g'%j['value'], 'e2e', '%.3g'%j['e2e']['value'], 'frac', round(j['roofline']['frac'],3), 'cpu', j['cpu_baseline'] and '%.3g'%j['cpu_baseline']['value'], 'launches', j['gpu_launches'])"
done
