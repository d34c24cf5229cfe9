"""GPU: a whole `LanPaint KSampler` run through the node API (ComfyUI replaced by minicomfy)
against the oracle's restatement of the same run (euler, karras-20, N=5) with the same seed."""
import pytest
import torch

import minicomfy
from _support import max_rel
from oracle import langevin_oracle as O

pytestmark = pytest.mark.gpu


def _denoiser(x, sigma, cond):
    """pointwise cond-dependent x0 prediction; `cond` is a float standing in for CONDITIONING"""
    return 0.7 * x + 0.1 * torch.tanh(x) + cond


class _OracleGuider:
    """What a patched CFGGuider returns to the engine: (x0 at cfg, x0 at cfg_BIG) (nodes.py:161-175)."""

    def __init__(self, pos, neg, cfg, cfg_big):
        self.inner_model = self
        self.model_sampling = O.VESampling()
        self.pos, self.neg, self.cfg, self.cfg_big = pos, neg, cfg, cfg_big
        self.calls = 0

    def __call__(self, x, sigma, model_options=None, seed=None):
        self.calls += 1
        c, u = _denoiser(x, sigma, self.pos), _denoiser(x, sigma, self.neg)
        return u + (c - u) * self.cfg, u + (c - u) * self.cfg_big


@pytest.mark.parametrize("mode,cfg_big", [("Image First", 5.0), ("Prompt First", -0.5)])
def test_ksampler_node_matches_oracle_run(mode, cfg_big, cuda_device):
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    dev = cuda_device
    g = torch.Generator().manual_seed(0)
    y = torch.randn(1, 4, 32, 32, generator=g)
    noise_mask = (torch.rand(1, 1, 32, 32, generator=g) < 0.5).float()   # 1 = regenerate
    patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(_denoiser), dev)
    seed, cfg = 3, 5.0
    (out,) = N.LanPaint_KSampler().sample(patcher, seed, 20, cfg, "euler", "karras", 0.3, -0.2,
                                          {"samples": y, "noise_mask": noise_mask}, 1.0, 5, mode, "", N.IMAGE_MODE)
    assert out["samples"].device.type == "cpu" and "noise_mask" in out
    eng = N.LAST_ENGINE["engine"]
    assert eng.substeps_done == 53 and eng.model_calls == 73
    assert patcher.pre_runs == 1 and patcher.cleanups == 1
    assert patcher.LanPaint_cfg_BIG == cfg_big and patcher.LanPaint_MinStepFrac == 1.0

    # the reference's run, restated: same CPU noise image, same CUDA generator stream for the Langevin draws
    noise = minicomfy.prepare_noise(y, seed)                       # seeds every device, like ComfyUI does
    model = _OracleGuider(0.3, -0.2, cfg, cfg_big)
    sig = O.karras_sigmas(20).to(dev)
    want = O.euler_inpaint(model, y.to(dev), noise.to(dev), noise_mask.expand(1, 4, 32, 32).to(dev), sig,
                           O.Hyper(n_steps=5, min_step_frac=1.0), max_denoise=True)
    assert model.calls == 73
    err = max_rel(out["samples"], want)
    assert err <= 1e-4, err


def test_custom_advanced_node_with_host_noise_object(cuda_device):
    """LanPaint_SamplerCustomAdvanced: NOISE object + GUIDER + SAMPLER + SIGMAS (the bench's e2e path)."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    dev = cuda_device
    g = torch.Generator().manual_seed(1)
    y = torch.randn(2, 4, 16, 16, generator=g)
    fixed_noise = torch.randn(2, 4, 16, 16, generator=g)
    noise_mask = (torch.rand(2, 1, 16, 16, generator=g) < 0.5).float()

    class HostNoise:
        seed = 11

        def generate_noise(self, latent):
            return fixed_noise

    patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(_denoiser), dev)
    guider = minicomfy.CFGGuider(patcher)
    guider.set_conds(0.3, -0.2)
    guider.set_cfg(4.0)
    sig = minicomfy.get_sigmas_karras(8, 0.0292, 14.6146)
    torch.manual_seed(5)
    out, den = N.LanPaint_SamplerCustomAdvanced().sample(HostNoise(), guider, minicomfy.ksampler("euler"), sig,
                                                         {"samples": y, "noise_mask": noise_mask}, 3, 5.0, 0.2,
                                                         "Image First")
    assert out["samples"].shape == y.shape and torch.isfinite(out["samples"]).all()
    assert den["samples"].shape == y.shape
    # batch of 2 with a [B] sigma: the reference's flow-form replace quirk applies (lanpaint.py:87-92)
    model = _OracleGuider(0.3, -0.2, 4.0, 4.0)
    torch.manual_seed(5)
    want = O.euler_inpaint(model, y.to(dev), fixed_noise.to(dev), noise_mask.expand(2, 4, 16, 16).to(dev), sig.to(dev),
                           O.Hyper(n_steps=3, lam=5.0, step_size=0.2, min_step_frac=1.0), max_denoise=True)
    assert max_rel(out["samples"], want) <= 1e-4


def test_heun_sampler_calls_wrapper_twice_per_step(cuda_device):
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    dev = cuda_device
    y = torch.randn(1, 4, 16, 16)
    noise_mask = (torch.rand(1, 1, 16, 16) < 0.5).float()
    patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(_denoiser), dev)
    (out,) = N.LanPaint_KSamplerAdvanced().sample(patcher, "enable", 1, 6, 5.0, "heun", "karras", 0.3, -0.2,
                                                  {"samples": y, "noise_mask": noise_mask}, 0, 10000, "disable",
                                                  2, 5.0, 0.2, "Image First", "", N.IMAGE_MODE)
    assert torch.isfinite(out["samples"]).all()
    assert N.LAST_ENGINE["engine"].model_calls > 6


def test_fused_cfg_equals_materialised_combines(cuda_device):
    """cond/uncond fed straight to lp_substep_cfg_f32 == the two eager cfg_function
    combines of nodes.py:175, bit for bit (same three roundings), and fewer launches."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    dev = cuda_device
    g = torch.Generator().manual_seed(0)
    y = torch.randn(2, 4, 32, 32, generator=g)
    noise_mask = (torch.rand(2, 1, 32, 32, generator=g) < 0.5).float()
    res = {}
    for fused in (True, False):
        patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(_denoiser), dev)
        patcher.model_options["lanpaint_b200"] = {"fused_cfg": fused}
        (out,) = N.LanPaint_KSampler().sample(patcher, 7, 12, 6.5, "euler", "karras", 0.3, -0.2,
                                              {"samples": y, "noise_mask": noise_mask}, 1.0, 4, "Prompt First", "",
                                              N.IMAGE_MODE)
        res[fused] = out["samples"]
    assert torch.equal(res[True], res[False])


def test_post_cfg_hook_disables_the_fusion(cuda_device):
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    from lanpaint_b200.engine import CfgPair
    dev = cuda_device
    x = torch.randn(1, 4, 8, 8, device=dev)
    model = minicomfy.BaseModel(_denoiser)
    t = torch.ones(1, device=dev)
    plain = N.sampling_function_LanPaint(model, x, t, -0.2, 0.3, 5.0, 2.0, model_options={})
    assert isinstance(plain, CfgPair)
    h0, h1 = plain.heads()
    hooked = N.sampling_function_LanPaint(model, x, t, -0.2, 0.3, 5.0, 2.0,
                                          model_options={"sampler_post_cfg_function": [lambda a: a["denoised"]]})
    assert isinstance(hooked, tuple) and torch.equal(hooked[0], h0) and torch.equal(hooked[1], h1)
    cfg1 = N.sampling_function_LanPaint(model, x, t, -0.2, 0.3, 1.0, 1.0, model_options={})
    assert isinstance(cfg1, tuple)   # uncond skipped at cfg == 1: nothing to fuse


def test_nodes_run_under_inference_mode(cuda_device):
    """ComfyUI executes nodes under torch.inference_mode(): inference tensors have no version counter and
    cannot be mutated outside it -- the caches and the in-place contract must cope."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    dev = cuda_device
    g = torch.Generator().manual_seed(2)
    y = torch.randn(1, 4, 32, 32, generator=g)
    noise_mask = (torch.rand(1, 1, 32, 32, generator=g) < 0.5).float()

    def run():
        patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(_denoiser), dev)
        (out,) = N.LanPaint_KSampler().sample(patcher, 5, 8, 4.0, "euler", "karras", 0.3, -0.2,
                                              {"samples": y, "noise_mask": noise_mask}, 1.0, 3, "Image First", "",
                                              N.IMAGE_MODE)
        return out["samples"]

    plain = run()
    with torch.inference_mode():
        inf = run()
    assert torch.equal(plain, inf.clone())


def test_flux_type_model_runs_on_the_flow_schedule(cuda_device):
    """ModelType.FLUX: rectified-flow change of variables, cfg_BIG forced to 1.0 (nodes.py:334-337), flow-form
    noise_scaling and inverse_noise_scaling; checked against the oracle's restatement of the same run."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    dev = cuda_device
    g = torch.Generator().manual_seed(4)
    y = torch.randn(1, 16, 16, 16, generator=g)
    noise_mask = (torch.rand(1, 1, 16, 16, generator=g) < 0.5).float()
    patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(_denoiser, model_type=minicomfy.ModelType.FLUX,
                                                         latent_channels=16), dev)
    sig = torch.tensor([0.98, 0.9, 0.75, 0.6, 0.45, 0.3, 0.15, 0.0])
    guider = minicomfy.CFGGuider(patcher)
    guider.set_conds(0.3, -0.2)
    guider.set_cfg(3.5)
    fixed_noise = torch.randn(1, 16, 16, 16, generator=g)

    class HostNoise:
        seed = 3

        def generate_noise(self, latent):
            return fixed_noise

    torch.manual_seed(9)
    out, _ = N.LanPaint_SamplerCustomAdvanced().sample(HostNoise(), guider, minicomfy.ksampler("euler"), sig,
                                                       {"samples": y, "noise_mask": noise_mask}, 4, 5.0, 0.2,
                                                       "Prompt First")
    eng = N.LAST_ENGINE["engine"]
    assert eng.IS_FLUX and guider.cfg_BIG == 1.0

    class FlowGuider(_OracleGuider):
        def __init__(self):
            super().__init__(0.3, -0.2, 3.5, 1.0)
            self.model_sampling = O.FlowSampling()
    model = FlowGuider()
    torch.manual_seed(9)
    want = O.euler_inpaint(model, y.to(dev), fixed_noise.to(dev), noise_mask.expand(1, 16, 16, 16).to(dev), sig.to(dev),
                           O.Hyper(n_steps=4, lam=5.0, step_size=0.2, min_step_frac=1.0, flow=True), max_denoise=True)
    assert eng.model_calls == model.calls
    assert max_rel(out["samples"], want) <= 1e-4


def test_node_path_with_graph_replay_equals_eager(cuda_device):
    """model_options["lanpaint_b200"] = {"cuda_graph": True}: the whole outer step, including the guider's
    cond/uncond evaluations, is captured once per sub-step count and replayed; results must not change."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    dev = cuda_device
    g = torch.Generator().manual_seed(6)
    y = torch.randn(1, 4, 32, 32, generator=g)
    noise_mask = (torch.rand(1, 1, 32, 32, generator=g) < 0.5).float()
    res = {}
    for graph in (False, True):
        patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(_denoiser), dev)
        patcher.model_options["lanpaint_b200"] = {"cuda_graph": graph}
        (out,) = N.LanPaint_KSampler().sample(patcher, 11, 20, 5.0, "euler", "karras", 0.3, -0.2,
                                              {"samples": y, "noise_mask": noise_mask}, 1.0, 5, "Image First", "",
                                              N.IMAGE_MODE)
        res[graph] = (out["samples"], N.LAST_ENGINE["engine"])
    assert torch.equal(res[True][0], res[False][0])
    eng = res[True][1]
    assert len([v for v in eng._graphs.values() if v]) == 6       # sub-step counts 5,4,3,2,1,0 of karras-20 x N=5
    assert eng.model_calls == 73 and eng.substeps_done == 53
