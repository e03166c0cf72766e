"""bench.py prints ONE JSON line with the keys the driver reads (both arms)."""
import json
import os
import subprocess
import sys

import pytest

from conftest import ROOT

BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def _run(args, timeout=600):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True,
                         timeout=timeout, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stdout[-2000:]
    return json.loads(lines[0])


def test_reference_arm_line(built_lib, monkeypatch):
    d = _run(["--impl", "reference", "--steps", "1", "--warmup", "0"])
    assert d["impl"] == "reference" and BASE_KEYS <= set(d)
    assert d["unit"] == "poses/s" and d["higher_is_better"] is True and d["vs_baseline"] is None
    # the arm runs the UNMODIFIED reference staged under oracle/_ref (kind "reference"), all host threads
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] and "oracle/_ref" in d["cpu_baseline"]["sample"]
    assert d["config"]["preset"] == "c2" and d["config"]["poses_per_gpu"] == 262144
    assert abs(d["ms_per_step"] - 1e3 * 262144 / d["value"]) < 1e-6 * d["ms_per_step"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


@pytest.mark.gpu
def test_gpu_arm_line_small_run():
    d = _run(["--poses", "4096", "--oil-steps", "20", "--steps", "1", "--warmup", "1", "--no-cpu"])
    assert BASE_KEYS | {"roofline", "clocks"} <= set(d)
    assert d["n_gpus"] == 1 and d["value"] > 0 and d["e2e"]["value"] > 0 and d["results_finite"] is True
    assert d["sharded_equals_unsharded_slice"] is True
    assert d["gpu_launches"] >= 20  # at least one launch per OIL step
    r = d["roofline"]
    assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and 0 < r["frac"] < 1.2 and r["achieved"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] == 4096 * (17 * 3 * 4 + 8 + 4)
