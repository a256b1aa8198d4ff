"""bench.py's reference arm runs on the CPU: its one JSON line must carry the contract's keys (the GPU arm prints the
same line plus `roofline`, `clocks` and a measured `cpu_baseline`)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    if not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libcmfrec_ref_f64.so")):
        pytest.skip("oracle/_ref is built where /root/reference exists (make -C oracle ref)")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                          "cfg1_explicit_cg_k16_f64", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["metric"] == "rows_solved_per_sec" and j["unit"] == "rows/s"
    assert j["higher_is_better"] is True and j["n_gpus"] == 1 and j["steps"] == 1 and j["value"] > 0
    assert j["config"]["workload"] == "cfg1_explicit_cg_k16_f64"
    assert j["cpu_baseline"]["kind"] == "reference" and j["cpu_baseline"]["cores"] >= 1 and j["cpu_baseline"]["sample"]
    assert j["e2e"] == dict(value=j["value"], unit="rows/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0)


def test_roofline_traffic_comes_from_the_committed_capture():
    sys.path.insert(0, ROOT)
    import bench
    traffic, src = bench.measured_dram_traffic(bench.WORKLOADS["ml10m_explicit_cg_k64_f32"])
    assert src == os.path.join("profiles", "r1_ncu_full_cg_sweep_ml10m.csv")
    assert 0.05 < traffic < 0.5      # GB per launch: the factors are served from L2, DRAM sees the CSR stream
    assert bench.measured_dram_traffic(bench.WORKLOADS["ml10m_explicit_chol_k64_f32"]) == (None, None)
