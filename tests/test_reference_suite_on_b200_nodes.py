"""The REFERENCE's own unit tests for the node layer, run against `lanpaint_b200/comfy_nodes.py`.

Where /root/reference exists (the build container), its test files are run unmodified, from where they lie, in a
subprocess whose import system hands them this repository's node module wherever they import the reference's
(`tests/_reference_suite_plugin.py`).  They cover b1 of SURVEY 8b the way the reference's CI does: the widget
surface and the retired hidden inputs (tests/test_node_params.py), the value sanitiser, the MinStepFrac inner-step
ramp (tests/test_min_step_frac.py), reshape_mask / prepare_mask incl. the video temporal union
(tests/test_reshape_mask.py), MiniMax-H3 AV-pack detection and the guarded optional imports
(tests/test_av_schedule.py; its numeric tests drive the reference ENGINE's internals and are left to the reference).
Also here: the package imports and lists its nodes with no ComfyUI at all, as node-diff CI needs (reference
tests/test_LanPaint.py, __init__.py:14-98)."""
import json
import os
import re
import subprocess
import sys

import pytest

from conftest import GOLDEN_DIR, ROOT

REF = "/root/reference"
FILES = ["test_node_params.py", "test_min_step_frac.py", "test_reshape_mask.py", "test_av_schedule.py"]
ENGINE_INTERNALS = ("audio_rows or without_audio or replace_step or score_model or add_none_dims or prepare_step_size")


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests")), reason="/root/reference exists only in the build container")
def test_reference_node_tests_pass_on_this_node_module(tmp_path):
    env = dict(os.environ, B200_ROOT=ROOT, PYTHONPATH=os.path.join(ROOT, "tests"))
    cmd = [sys.executable, "-m", "pytest", "-p", "_reference_suite_plugin", "-p", "no:cacheprovider", "-q", "-W", "ignore",
           "--rootdir", str(tmp_path), "-k", f"not ({ENGINE_INTERNALS})"] + [os.path.join(REF, "tests", f) for f in FILES]
    res = subprocess.run(cmd, capture_output=True, text=True, cwd=str(tmp_path), env=env, timeout=600)
    tail = res.stdout[-3000:] + res.stderr[-2000:]
    assert res.returncode == 0, tail
    m = re.search(r"(\d+) passed", res.stdout)
    assert m and int(m.group(1)) >= 19, tail
    assert "failed" not in res.stdout and "error" not in res.stdout.lower().replace("errors", ""), tail
    loaded = re.search(r"b200-alias: node module loaded from (\S+) x(\d+)", res.stdout)
    assert loaded and loaded.group(1).endswith("lanpaint_b200/comfy_nodes.py") and int(loaded.group(2)) >= 5, tail


def test_package_lists_its_nodes_without_comfyui(tmp_path):
    """reference tests/test_LanPaint.py + __init__.py:90-98: importable, NODE_CLASS_MAPPINGS introspectable, where
    neither ComfyUI nor any stand-in is installed (a clean interpreter, not this test process)."""
    code = (
        "import sys, json\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "import lanpaint_b200\n"
        "m = lanpaint_b200.NODE_CLASS_MAPPINGS\n"
        "import comfy\n"
        "assert getattr(comfy, '__lanpaint_b200_tooling_stub__', False)\n"
        "sched = json.dumps(comfy.samplers.KSampler.SCHEDULERS)\n"
        "out = {k: json.loads(json.dumps(v.INPUT_TYPES(), ensure_ascii=False).replace(sched, json.dumps(['<SCHEDULERS>'])))"
        " for k, v in m.items()}\n"
        "print(json.dumps({'inputs': out, 'names': lanpaint_b200.NODE_DISPLAY_NAME_MAPPINGS, 'web': lanpaint_b200.WEB_DIRECTORY}))\n")
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd=str(tmp_path), timeout=300)
    assert res.returncode == 0, res.stderr[-2000:]
    got = json.loads(res.stdout.strip().splitlines()[-1])
    api = json.load(open(os.path.join(GOLDEN_DIR, "node_api.json")))
    for name in ("LanPaint_KSampler", "LanPaint_KSamplerAdvanced", "LanPaint_SamplerCustom", "LanPaint_SamplerCustomAdvanced"):
        assert got["inputs"][name] == api[name]["INPUT_TYPES"], name
        assert got["names"][name] == api[name]["display_name"]
    assert got["web"] == "./web" and os.path.isdir(os.path.join(ROOT, "lanpaint_b200", "web"))
