"""bench.py's JSON contract, exercised on the CPU through the reference arm (the GPU arm prints the same keys
plus roofline/clocks; it is run by the driver on the B200)."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_one_contract_line():
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--ref-requests", "2", "--gpus", "1"], capture_output=True, text=True,
                         timeout=600)
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [l for l in res.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "sub-steps/s" and d["higher_is_better"] is True
    assert d["metric"].startswith("Langevin sub-steps/sec") and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and d["data"] == "synthetic" and d["n_gpus"] == 1 and d["steps"] == 1
    assert d["config"]["workload"].startswith("sdxl_1024_inpaint") and d["config"]["substeps_per_request"] == 53
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["value"] == d["value"] > 0 and "sample" in cb
    assert d["config"]["reference_requests"] == 2
    e2e = dict(d["e2e"])
    api = e2e.pop("api", None)
    assert e2e == {"value": d["value"], "unit": "sub-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if cb["kind"] == "reference":      # oracle/_ref built: the arm is the reference's own node call, like the GPU arm's
        assert "LanPaint_KSampler.sample" in api and "LanPaint_KSampler.sample" in cb["sample"]
        assert d["config"]["api"] == api


def test_non_zero_ranks_of_the_reference_arm_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--gpus", "2"], capture_output=True, text=True, timeout=120, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""
