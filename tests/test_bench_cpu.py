"""CPU checks of bench.py: the reference arm (`--impl reference`) runs on the host alone and prints the contract's JSON
line for every configuration; under a multi-rank launch only rank 0 works."""

import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "fedoo")


def _run(args, env=None):
    e = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    e.update(env or {})
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, env=e,
                         timeout=600, cwd=ROOT)  # fmt: skip
    assert out.returncode == 0, out.stderr[-2000:]
    return [json.loads(l) for l in out.stdout.splitlines() if l.startswith("{")]


@pytest.mark.parametrize("config,edge", [("hex8", 6), ("heat_tet4", 5), ("tet10", 3), ("j2_plate", 4)])
def test_reference_arm_line(config, edge):
    lines = _run(["--impl", "reference", "--config", config, "--cpu-n", str(edge), "--steps", "1", "--warmup", "1"])
    assert len(lines) == 1
    d = lines[0]
    assert d["impl"] == "reference" and d["unit"] == "Melem/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["value"] == d["value"] and cb["cores"] >= 1 and "elements" in cb["sample"]
    # the real reference wherever it exists and can run the configuration; the NumPy port for J2 (simcoon absent)
    expect = "reference" if (os.path.isdir(REF) and config != "j2_plate") else "port"
    assert cb["kind"] == expect
    assert "sample" in d["config"] and d["config"]["workload"]
    assert d["gpu_launches"] == 0 and d["dtype"] == "f64"


def test_reference_arm_other_ranks_exit_quietly():
    assert _run(["--impl", "reference", "--cpu-n", "4", "--steps", "1", "--warmup", "1"], env={"RANK": "3", "WORLD_SIZE": "8"}) == []
