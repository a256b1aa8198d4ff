#!/bin/bash
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -2
timeout 300 python tools/e2e_timing.py 2>&1 | tail -8
timeout 900 python bench.py 2>/dev/null | python -c "
import json,sys; j=json.loads(sys.stdin.read()); print(j['config']['workload'], 'ms', round(j['ms_per_step'],3), 'value', '%.3g'%j['value'], 'e2e', '%.3g'%j['e2e']['value'], 'frac', round(j['roofline']['frac'],3))"
