#!/bin/bash
# round-2 closing run on one GPU: the whole -m gpu suite, smoke(), both bench arms of the default workload, the other
# configurations' lines, and the ncu launch list of the default bench command
mkdir -p gpurun_out/r2final
O=gpurun_out/r2final
S=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 > $O/gputest.log; cat $O/gputest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 > $O/smoke.log; cat $O/smoke.log
timeout 600 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 600 python bench.py --impl reference > $O/bench_reference.json 2> $O/bench_reference.err
for w in lastfm_implicit_cg_k64_f32 ml10m_explicit_cg_k64_f32_implicit_features; do
  timeout 600 python bench.py --workload $w --no-cpu-baseline > $O/bench_$w.json 2> $O/bench_$w.err
done
timeout 600 python bench.py --workload ml10m_explicit_chol_k128_f64_sideinfo --no-cpu-baseline --steps 5 > $O/bench_ml10m_explicit_chol_k128_f64_sideinfo.json 2> $O/bench_cfg3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/ncu_launches_default.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > $O/ncu_bench.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2final/bench_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms %.3f" % d.get("ms_per_step", -1), "value %.4g" % d["value"], "e2e %.4g" % (d["e2e"]["value"] if d.get("e2e") else -1),
              "frac %.3f" % d["roofline"]["frac"] if d.get("roofline") else "", d.get("cpu_baseline", {}) and d["cpu_baseline"].get("value"))
    except Exception as e:
        print(f, "failed", e)
PY
echo "total $(( $(date +%s) - S )) s"
