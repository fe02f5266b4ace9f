"""bench.py prints exactly one JSON line with the keys the driver reads (task contract ④)."""
import json
import subprocess
import sys

import pytest

from conftest import ROOT

COMMON = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
          "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"}


def run_bench(*args):
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), *args], capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, lines
    return json.loads(lines[0])


def test_reference_arm_line():
    """`--impl reference`: the CPU restatement of the reference's simd path on a bounded tile sample."""
    d = run_bench("--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "1")
    assert COMMON <= set(d) and d["impl"] == "reference" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["metric"] == "weighted LD site-pairs/sec" and d["unit"] == "site-pairs/s" and d["higher_is_better"] is True


@pytest.mark.gpu
def test_our_arm_line():
    d = run_bench("--workload", "tiny", "--steps", "2", "--warmup", "3")
    assert COMMON | {"roofline", "clocks", "gpu_launches", "stages_ms", "hbm_stages"} <= set(d)
    r = d["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and r["bound"] == "tensor"
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert d["gpu_launches"] > 0 and d["value"] > 0 and d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 512 * 3000 and d["e2e"]["d2h_bytes_per_step"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] > 0
    assert d["config"]["workload"].startswith("synthetic") and "model" not in d["config"]
    assert set(d["clocks"]) >= {"sm_mhz", "sm_max_mhz", "reasons"}
    assert {"used", "candidates", "sample_pairs", "sample_candidates"} <= set(d["screen"])
    # both arms name the workload with the same dictionary (the driver compares them)
    ref = run_bench("--impl", "reference", "--workload", "tiny", "--steps", "1", "--warmup", "1")
    assert ref["config"] == d["config"] and "l2" in d["config"]
