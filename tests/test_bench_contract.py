"""bench.py's reference arm runs without a GPU: check the JSON line it prints against the driver's contract (the keys, the
metric / unit, the cpu_baseline and e2e objects of the reference arm).  The GPU arm's line is checked on the B200 box
(tests marked gpu)."""
import json
import os
import subprocess
import sys

import pytest

from harness import ROOT

REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config"]


def run_bench(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + list(args), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.strip().splitlines() if l.startswith("{")]
    assert len(lines) == 1, r.stdout[-2000:]          # ONE JSON line
    return json.loads(lines[0])


def test_reference_arm_line():
    d = run_bench("--impl", "reference", "--workload", "c1_16_pan", "--steps", "4", "--warmup", "3")
    for k in REQUIRED:
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "source-samples/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["vs_baseline"] is None and d["dtype"] == "f32" and d["data"] == "synthetic"
    assert d["config"]["workload"] == "c1_16_pan"
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_other_ranks_print_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--workload", "c1_16_pan", "--steps", "2"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


@pytest.mark.gpu
def test_gpu_arm_line():
    d = run_bench("--workload", "c2_4096_resample_pan", "--steps", "20", "--warmup", "3", "--no-also", "--no-full-parity", "--blocks", "3")
    for k in REQUIRED + ["e2e", "gpu_launches", "roofline", "clocks", "cpu_baseline", "parity"]:
        assert k in d, k
    assert d["n_gpus"] == 1 and d["gpu_launches"] > 0 and d["value"] > 0
    assert set(["value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "kernels_ms"]) <= set(d["e2e"]) and d["e2e"]["h2d_bytes_per_step"] > 0
    assert set(["bound", "achieved", "peak", "unit", "frac", "traffic"]) <= set(d["roofline"]) and d["roofline"]["bound"] == "hbm"
    assert set(["value", "unit", "cores", "kind", "sample"]) <= set(d["cpu_baseline"])
    assert set(["sm_mhz", "sm_max_mhz", "reasons"]) <= set(d["clocks"])
    assert d["parity"]["rel_rms"] <= 1e-5
    assert d["timing"]["blocks"] == 3 and len(d["timing"]["value_block_ms_per_step"]) == 3
    assert d["e2e"]["all_move"]["h2d_bytes_per_step"] >= 4096 * 16 and d["roofline"]["l2_resident"] is True and d["roofline"]["frac"] is None
    assert d["latency_us"]["p50"] > 0
