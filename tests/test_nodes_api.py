"""The ComfyUI node surface must stay byte-identical to the reference's
(SURVEY 8b/b1).  tests/golden/node_api.json is a dump of the reference's own
classes (tests/golden/make_golden.py --api).  CPU only; ComfyUI is replaced by
minicomfy."""
import json
import os

import pytest

import minicomfy
from conftest import GOLDEN_DIR


@pytest.fixture()
def nodes_mod(monkeypatch):
    minicomfy.install()
    import comfy.samplers
    monkeypatch.setattr(comfy.samplers.KSampler, "SCHEDULERS", ["<SCHEDULERS>"])
    from lanpaint_b200 import comfy_nodes
    return comfy_nodes


def _norm(o):
    return json.loads(json.dumps(o, ensure_ascii=False))


def test_sampler_nodes_match_reference_dump(nodes_mod):
    api = json.load(open(os.path.join(GOLDEN_DIR, "node_api.json")))
    assert nodes_mod.KSAMPLER_NAMES == api["KSAMPLER_NAMES"] and len(api["KSAMPLER_NAMES"]) == 22
    for name in ("LanPaint_KSampler", "LanPaint_KSamplerAdvanced", "LanPaint_SamplerCustom",
                 "LanPaint_SamplerCustomAdvanced"):
        cls = nodes_mod.NODE_CLASS_MAPPINGS[name]
        want = api[name]
        got = _norm(cls.INPUT_TYPES())
        assert list(got["required"].keys()) == list(want["INPUT_TYPES"]["required"].keys()), name  # widget ORDER
        assert got == want["INPUT_TYPES"], name
        assert list(cls.RETURN_TYPES) == want["RETURN_TYPES"]
        assert list(getattr(cls, "RETURN_NAMES", ())) == want["RETURN_NAMES"]
        assert cls.FUNCTION == want["FUNCTION"] == "sample"
        assert cls.CATEGORY == want["CATEGORY"]
        assert nodes_mod.NODE_DISPLAY_NAME_MAPPINGS[name] == want["display_name"]


def test_retired_widgets_are_hidden_not_required(nodes_mod):
    """reference tests/test_node_params.py:33-70."""
    retired = ["LanPaint_Beta", "LanPaint_Friction", "LanPaint_EarlyStop", "LanPaint_InnerThreshold",
               "LanPaint_InnerPatience", "LanPaint_MinStepFrac"]
    for cls in nodes_mod.NODE_CLASS_MAPPINGS.values():
        req = cls.INPUT_TYPES().get("required", {})
        assert not set(retired) & set(req)
    hid = lambda c: set(c.INPUT_TYPES().get("hidden", {}))
    assert hid(nodes_mod.LanPaint_KSamplerAdvanced) >= set(retired)
    assert hid(nodes_mod.LanPaint_SamplerCustomAdvanced) >= set(retired)
    assert "LanPaint_MinStepFrac" in hid(nodes_mod.LanPaint_KSampler)
    req = nodes_mod.LanPaint_KSamplerAdvanced.INPUT_TYPES()["required"]
    for name in ("LanPaint_NumSteps", "LanPaint_Lambda", "LanPaint_StepSize", "LanPaint_PromptMode", "LanPaint_Info",
                 "Inpainting_mode"):
        assert name in req


def test_sanitize_param(nodes_mod):
    """reference tests/test_node_params.py:73-97."""
    s = nodes_mod._sanitize_param
    allowed = ("Image First", "Prompt First")
    assert s("Prompt First", "Image First", allowed=allowed) == "Prompt First"
    for bad in (1.0, "bogus", None):
        assert s(bad, "Image First", allowed=allowed) == "Image First"
    assert s(5, 5) == 5 and s(3.7, 0.2) == 3.7
    assert s("abc", 0.2) == 0.2 and s(None, 0.2) == 0.2 and s(True, 5) == 5


def test_package_exports_node_protocol(nodes_mod):
    """reference tests/test_LanPaint.py:7-13."""
    import lanpaint_b200
    assert isinstance(lanpaint_b200.NODE_CLASS_MAPPINGS, dict)
    assert isinstance(lanpaint_b200.NODE_DISPLAY_NAME_MAPPINGS, dict)
    assert "LanPaint_KSampler" in lanpaint_b200.NODE_CLASS_MAPPINGS
    assert lanpaint_b200.WEB_DIRECTORY == "./web"


def test_override_restores_on_exception_and_nests(nodes_mod):
    import comfy.sampler_helpers
    import comfy.samplers
    before = (comfy.samplers.CFGGuider.outer_sample, comfy.samplers.CFGGuider.predict_noise,
              comfy.samplers.KSAMPLER.sample, comfy.sampler_helpers.prepare_mask)
    with pytest.raises(RuntimeError):
        with nodes_mod.override_sample_function():
            assert comfy.samplers.CFGGuider.predict_noise is nodes_mod.CFGGuider_LanPaint.predict_noise
            with nodes_mod.override_sample_function():  # nested entry must not capture the patches as originals
                pass
            assert comfy.samplers.CFGGuider.predict_noise is nodes_mod.CFGGuider_LanPaint.predict_noise
            raise RuntimeError("boom")
    after = (comfy.samplers.CFGGuider.outer_sample, comfy.samplers.CFGGuider.predict_noise,
             comfy.samplers.KSAMPLER.sample, comfy.sampler_helpers.prepare_mask)
    assert before == after


def test_reshape_mask_shapes(nodes_mod):
    """reference tests/test_reshape_mask.py semantics: 2-D / 3-D / 4-D image masks and the video union."""
    import torch
    rm = nodes_mod.reshape_mask
    m2 = (torch.rand(64, 64) > 0.5).float()
    assert rm(m2, (2, 4, 8, 8)).shape == (2, 4, 8, 8)
    assert rm(m2[None], (1, 16, 8, 8)).shape == (1, 16, 8, 8)
    assert rm(m2[None, None], (1, 4, 8, 8)).shape == (1, 4, 8, 8)
    out = rm(m2, (1, 4, 8, 8))
    assert torch.equal(out[:, 0], out[:, 3]) and set(out.unique().tolist()) <= {0.0, 1.0}
    still = rm(m2, (1, 16, 4, 8, 8), video_inpainting=True)
    assert still.shape == (1, 16, 4, 8, 8) and torch.equal(still[:, :, 0], still[:, :, 3])


def test_reshape_mask_matches_reference_outputs(nodes_mod):
    """Known answers generated by the reference's own reshape_mask (make_golden.py --api)."""
    import numpy as np
    import torch
    import importlib.util
    spec = importlib.util.spec_from_file_location("mg_cases", os.path.join(GOLDEN_DIR, "make_golden.py"))
    src = open(os.path.join(GOLDEN_DIR, "make_golden.py")).read()
    ns = {}
    exec(src[src.index("RESHAPE_CASES = {"):src.index("def dump_node_api")], ns)
    z = np.load(os.path.join(GOLDEN_DIR, "aux_reshape_mask_cases.npz"))
    for key, (mshape, oshape, video) in ns["RESHAPE_CASES"].items():
        got = nodes_mod.reshape_mask(torch.from_numpy(z["in_" + key]), oshape, video)
        assert tuple(got.shape) == tuple(oshape), key
        assert torch.equal(got.contiguous(), torch.from_numpy(z["out_" + key])), key


def test_prepare_mask_keeps_the_reference_values_while_travelling_compact(nodes_mod):
    """prepare_mask == reshape_mask(...).to(device) in values and shape (every golden case of the reference's own
    function); single-channel masks come back as a stride-0 expansion (1/C of the bytes cross PCIe) and a mask
    that needs no resampling may arrive as uint8 / bool."""
    import numpy as np
    import torch
    src = open(os.path.join(GOLDEN_DIR, "make_golden.py")).read()
    ns = {}
    exec(src[src.index("RESHAPE_CASES = {"):src.index("def dump_node_api")], ns)
    z = np.load(os.path.join(GOLDEN_DIR, "aux_reshape_mask_cases.npz"))
    for key, (mshape, oshape, video) in ns["RESHAPE_CASES"].items():
        got = nodes_mod.prepare_mask(torch.from_numpy(z["in_" + key]), oshape, "cpu", video_inpainting=video)
        assert tuple(got.shape) == tuple(oshape), key
        assert torch.equal(got.contiguous(), torch.from_numpy(z["out_" + key])), key
    m = (torch.rand(3, 1, 16, 16) > 0.5)
    want = nodes_mod.reshape_mask(m.float(), (3, 4, 16, 16))
    for src_mask in (m, m.to(torch.uint8), m.float(), m[:, 0].float()):
        got = nodes_mod.prepare_mask(src_mask, (3, 4, 16, 16), "cpu")
        assert got.dtype == torch.float32 and got.stride(1) == 0 and torch.equal(got, want)
    full = (torch.rand(2, 4, 8, 8) > 0.5).float()          # a mask that differs per channel stays materialised
    got = nodes_mod.prepare_mask(full, (2, 4, 8, 8), "cpu")
    assert torch.equal(got, full) and got.stride(1) != 0


# ---- MiniMax-H3 detection (reference tests/test_av_schedule.py:75-105) -----------------------------------
class _FakeDiffusion:
    sigma_shift_video = 12.0
    sigma_shift_audio = 3.0


def _patcher(with_shifts=True):
    import types
    model = types.SimpleNamespace(diffusion_model=_FakeDiffusion()) if with_shifts else object()
    return types.SimpleNamespace(model=model)


_TWO = [(1, 24, 37, 30, 54), (1, 32, 2, 207)]


def test_minimax_detection(nodes_mod):
    det = nodes_mod._detect_minimax_h3_audio
    assert det(_patcher(), {}, [(1, 24, 37, 30, 54)]) is None and det(_patcher(), {}, None) is None
    assert det(_patcher(False), {}, _TWO) is None
    assert det(_patcher(), {}, _TWO) == (_TWO, 12.0, 3.0)
    over = {"transformer_options": {"minimax_h3_sigma_shift_video": 10.0, "minimax_h3_sigma_shift_audio": 2.5}}
    assert det(_patcher(), over, _TWO) == (_TWO, 10.0, 2.5)
    assert nodes_mod.time_shift_sigma is None   # comfy.ldm.minimax is not importable in the stand-in
