"""bench.py's stdout contract: ONE compact JSON line (round 1 printed 20 kB and the driver's tail window lost it), with
the keys the driver and the judge read.  The reference arm runs on CPU everywhere; the repo's own arm needs a B200."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e"}


def _run(*flags):
    p = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *flags], cwd=ROOT, capture_output=True, text=True,
                       timeout=900)
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines                       # nothing but the JSON line on stdout
    assert len(lines[0]) < 4096, len(lines[0])          # fits any sane tail window
    return json.loads(lines[0])


def _check_base(line):
    assert BASE_KEYS <= set(line), BASE_KEYS - set(line)
    assert line["metric"] == "audio_seconds_per_second_encode_decode" and line["unit"] == "audio-s/s"
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["higher_is_better"] is True
    assert line["scaling"] == "weak" and line["vs_baseline"] is None and line["data"] == "synthetic"
    assert "workload" in line["config"] and "model" not in line["config"]
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])


def test_reference_arm_prints_one_parseable_line():
    line = _run("--impl", "reference", "--steps", "1", "--warmup", "1", "--cpu-seconds", "1.0")
    _check_base(line)
    assert line["impl"] == "reference" and line["dtype"] == "fp32"
    assert line["steps"] == 1 and line["warmup"] == 1 and line["n_gpus"] == 1
    cb = line["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == line["value"]
    assert line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert "cpu_sample" in line["config"]               # says what the bounded sample was, next to the workload it stands for


@pytest.mark.gpu
def test_own_arm_prints_one_parseable_line():
    line = _run("--steps", "2", "--warmup", "3", "--clips", "8", "--no-cpu-baseline")
    _check_base(line)
    assert "impl" not in line or line["impl"] != "reference"
    assert line["dtype"] == "bf16" and line["steps"] == 2 and line["warmup"] == 3
    assert line["gpu_launches"] > 0
    assert line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert 0.3 * line["value"] < line["e2e"]["value"] < 2.0 * line["value"]   # its own measurement, same order as `value`
    r = line["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r)
    assert r["bound"] in ("hbm", "tensor") and 0 < r["frac"] < 1.1 and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-3
    c = line["clocks"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(c) and isinstance(c["reasons"], list)
    if c["sm_mhz"] is not None:                         # a 2-step run of 8 clips may end before the first nvidia-smi sample
        assert 0 < c["sm_mhz"] <= c["sm_max_mhz"]
    assert len(line["per_rank_ms"]) == 1
