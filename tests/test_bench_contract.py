"""bench.py contract checks that run without a GPU: the reference arm end to end (it times the unmodified reference's
`-c` path), and the JSON schema of the last committed B200 bench line."""
import json
import subprocess
import sys

import pytest

from _refio import REPO

REQUIRED = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "e2e"}


@pytest.mark.skipif(not (REPO / "oracle" / "_ref" / "cudaSaTabsearch_ref").exists(), reason="reference binary not built")
def test_reference_arm_prints_one_valid_line():
    p = subprocess.run([sys.executable, str(REPO / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, cwd=str(REPO))
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [ln for ln in p.stdout.split("\n") if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert REQUIRED <= set(d) and d["impl"] == "reference" and d["unit"] == "structures/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_committed_bench_line_has_every_contract_key():
    d = json.loads((REPO / "profiles" / "r02z_bench.json").read_text())
    assert REQUIRED | {"gpu_launches", "clocks", "roofline", "cpu_baseline"} <= set(d)
    assert d["higher_is_better"] is True and d["scaling"] in ("weak", "strong") and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["gpu_launches"] > 0 and d["data"] == "synthetic"
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(d["roofline"])
    assert abs(d["roofline"]["frac"] - d["roofline"]["achieved"] / d["roofline"]["peak"]) < 1e-9
    assert {"value", "unit", "cores", "kind", "sample"} <= set(d["cpu_baseline"])
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"])
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
