#!/bin/bash
# last run of the round on one GPU after the final code changes: whole -m gpu suite, smoke(), default bench, config 4
mkdir -p gpurun_out/r2final
O=gpurun_out/r2final
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/gputest.log; cat $O/gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 > $O/smoke.log; cat $O/smoke.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --workload ml10m_explicit_cg_k64_f32_implicit_features --no-cpu-baseline > $O/bench_ml10m_explicit_cg_k64_f32_implicit_features.json 2> $O/bench_cfg4.err
python - <<'PY'
import json
for f in ("bench_default", "bench_ml10m_explicit_cg_k64_f32_implicit_features"):
    try:
        d = json.loads([l for l in open("gpurun_out/r2final/%s.json" % f) if l.startswith("{")][-1])
        print(f, "ms %.3f" % d["ms_per_step"], "value %.4g" % d["value"], "e2e %.4g" % d["e2e"]["value"], "frac %.3f" % d["roofline"]["frac"])
    except Exception as e:
        print(f, "failed", e)
PY
