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
    # and the graph-replayed path entirely under inference mode (static buffers, capture and replay all created there)
    N._ENGINES.clear()
    patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(_denoiser), dev)
    with torch.inference_mode():
        outs, modes = [], []
        for _ in range(4):
            (out,) = N.LanPaint_KSampler().sample(patcher, 5, 8, 4.0, "euler", "karras", 0.3, -0.2,
                                                  {"samples": y, "noise_mask": noise_mask}, 1.0, 3, "Image First", "",
                                                  N.IMAGE_MODE)
            outs.append(out["samples"].clone())
            modes.append(N.LAST_RUN["mode"])
    assert modes[0] == "eager" and modes[-1] in ("steps", "job"), modes
    for o in outs:
        assert torch.equal(o, plain)


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


def _ksampler_run(N, dev, y, noise_mask, opts=None, seed=11, cfg=5.0, sampler="euler", steps=20, n=5,
                  patcher=None, conds=(0.3, -0.2)):
    if patcher is None:
        patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(_denoiser), dev)
    if opts is not None:
        patcher.model_options["lanpaint_b200"] = opts
    (out,) = N.LanPaint_KSampler().sample(patcher, seed, steps, cfg, sampler, "karras", conds[0], conds[1],
                                          {"samples": y, "noise_mask": noise_mask}, 1.0, n, "Image First", "",
                                          N.IMAGE_MODE)
    return out["samples"], patcher


def test_node_default_is_graph_replay_and_matches_oracle(cuda_device):
    """The node path with NO options (what a workflow gets): Euler is run by the host-owned fused loop; the first
    job of a configuration launches eagerly, the second captures one CUDA graph per outer step (ComfyUI always
    passes a progress callback), later ones only replay.  Every one of them must equal the oracle's restatement
    of the reference run with the same seed (rng="torch": the reference's own randn stream), and each other bit
    for bit."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    N._ENGINES.clear()
    dev = cuda_device
    g = torch.Generator().manual_seed(6)
    y = torch.randn(2, 4, 32, 32, generator=g)
    noise_mask = (torch.rand(2, 1, 32, 32, generator=g) < 0.5).float()
    outs = []
    for deferred, want_modes, want_captures in ((False, ["eager", "steps", "steps", "steps"], 20),
                                                (True, ["eager", "job", "job", "job"], 1)):
        # a job this short (a pointwise network) runs as ONE graph with the callbacks delivered right after it;
        # {"deferred_callbacks": False} -- or a job longer than 50 ms, i.e. any real network -- keeps one graph per
        # outer step with the callback between them
        patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(_denoiser), dev)
        modes = []
        calls0 = minicomfy.PROGRESS["calls"]
        for _ in range(4):
            o, _ = _ksampler_run(N, dev, y, noise_mask, patcher=patcher, opts={"deferred_callbacks": deferred})
            outs.append(o)
            modes.append(N.LAST_RUN["mode"])
            eng = N.LAST_ENGINE["engine"]
            assert eng.model_calls == 73 and eng.substeps_done == 53 and eng.rng == "torch"
        assert modes == want_modes, modes
        assert minicomfy.PROGRESS["calls"] - calls0 == 4 * 20 and minicomfy.PROGRESS["last"] == (20, 20)
        assert N.LAST_RUN["job"].captures == want_captures
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    noise = minicomfy.prepare_noise(y, 11)        # re-seeds every generator exactly like the node call did
    model = _OracleGuider(0.3, -0.2, 5.0, 5.0)
    want = O.euler_inpaint(model, y.to(dev), noise.to(dev), noise_mask.expand(2, 4, 32, 32).to(dev),
                           O.karras_sigmas(20).to(dev), O.Hyper(n_steps=5, min_step_frac=1.0), max_denoise=True)
    assert max_rel(outs[0], want) <= 1e-4
    # the un-fused route (k-diffusion's own loop around the per-sigma wrapper, plain launches) agrees too
    plain, _ = _ksampler_run(N, dev, y, noise_mask, opts={"cuda_graph": False, "fused_sampler": False})
    assert N.LAST_RUN["fused"] is False
    assert max_rel(outs[0], plain) <= 2e-5


def test_node_cache_is_keyed_on_what_a_graph_bakes_in(cuda_device):
    """A different guidance scale, prompt object or seed must never replay a stale graph."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    N._ENGINES.clear()
    dev = cuda_device
    g = torch.Generator().manual_seed(8)
    y = torch.randn(1, 4, 32, 32, generator=g)
    noise_mask = (torch.rand(1, 1, 32, 32, generator=g) < 0.5).float()
    patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(_denoiser), dev)
    for _ in range(5):   # cfg 5 is now graph-replayed (eager -> [per-step graphs ->] whole-job graph once it proved short)
        a5, _ = _ksampler_run(N, dev, y, noise_mask, patcher=patcher, steps=8, n=3)
    assert N.LAST_RUN["mode"] == "job"
    a7, _ = _ksampler_run(N, dev, y, noise_mask, patcher=patcher, steps=8, n=3, cfg=7.0)
    assert N.LAST_RUN["mode"] == "eager"
    ref7, _ = _ksampler_run(N, dev, y, noise_mask, opts={"cuda_graph": False}, steps=8, n=3, cfg=7.0)
    assert torch.equal(a7, ref7) and not torch.equal(a7, a5)
    # another seed replays the same graphs with a fresh noise image and a fresh randn stream
    b5, _ = _ksampler_run(N, dev, y, noise_mask, patcher=patcher, steps=8, n=3, seed=12)
    assert N.LAST_RUN["mode"] == "job"
    refb, _ = _ksampler_run(N, dev, y, noise_mask, opts={"cuda_graph": False}, steps=8, n=3, seed=12)
    assert torch.equal(b5, refb) and not torch.equal(b5, a5)
    # a float conditioning of equal value is the same prompt for minicomfy; a different value is not
    c5, _ = _ksampler_run(N, dev, y, noise_mask, patcher=patcher, steps=8, n=3, conds=(0.5, -0.2))
    refc, _ = _ksampler_run(N, dev, y, noise_mask, opts={"cuda_graph": False}, steps=8, n=3, conds=(0.5, -0.2))
    assert torch.equal(c5, refc)


def test_whole_job_graph_when_nobody_wants_a_callback(cuda_device):
    """guider.sample(...) without a callback (custom hosts): the fused loop is ONE graph for the whole job."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    N._ENGINES.clear()
    dev = cuda_device
    g = torch.Generator().manual_seed(3)
    y = torch.randn(1, 4, 32, 32, generator=g)
    noise = torch.randn(1, 4, 32, 32, generator=g)
    noise_mask = (torch.rand(1, 1, 32, 32, generator=g) < 0.5).float()
    patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(_denoiser), dev)
    N._set_hyper(patcher, num_steps=4, cfg=4.0, prompt_mode="Image First")
    sig = minicomfy.get_sigmas_karras(10, 0.0292, 14.6146)
    outs, modes = [], []
    for _ in range(3):
        guider = minicomfy.CFGGuider(patcher)
        guider.set_conds(0.3, -0.2)
        guider.set_cfg(4.0)
        torch.manual_seed(21)
        with N.override_sample_function():
            outs.append(guider.sample(noise, y, minicomfy.ksampler("euler"), sig, denoise_mask=noise_mask, seed=1))
        modes.append(N.LAST_RUN["mode"])
    assert modes == ["eager", "job", "job"]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert N.LAST_RUN["job"].captures == 1


def test_other_samplers_replay_one_graph_per_outer_step(cuda_device):
    """heun (any sampler that is not plain Euler) goes through k-diffusion's own loop; the per-sigma wrapper
    replays one CUDA graph per outer step, cached across sample() calls.  Same results as plain launches."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    N._ENGINES.clear()
    dev = cuda_device
    g = torch.Generator().manual_seed(5)
    y = torch.randn(1, 4, 32, 32, generator=g)
    noise_mask = (torch.rand(1, 1, 32, 32, generator=g) < 0.5).float()
    plain, _ = _ksampler_run(N, dev, y, noise_mask, opts={"cuda_graph": False}, sampler="heun", steps=10, n=3)
    calls = N.LAST_ENGINE["engine"].model_calls
    patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(_denoiser), dev)
    for k in range(3):
        got, _ = _ksampler_run(N, dev, y, noise_mask, patcher=patcher, sampler="heun", steps=10, n=3,
                               opts={"sampler_graph": False})
        assert N.LAST_RUN["fused"] is False
        assert torch.equal(got, plain), k
        assert N.LAST_ENGINE["engine"].model_calls == calls
    eng = N.LAST_ENGINE["engine"]
    assert len([v for v in eng._graphs.values() if v]) >= 3


def test_av_flat_pack_through_the_node_layer(cuda_device, monkeypatch):
    """MiniMax-H3 AV pack entered through the nodes (PackedMask from the per-sigma wrapper, audio rows on their
    own schedule) == the engine seam driven with tensor masks and the same per-step arguments."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    from lanpaint_b200.engine import LanPaint
    from lanpaint_b200.schedule import times_from_sigma
    N._ENGINES.clear()
    dev = cuda_device

    def shift_sigma(s, sv, sa):       # stand-ins for comfy.ldm.minimax.model.time_shift_sigma / _slope
        return sa * s / (sv + (sa - sv) * s)

    def shift_slope(s, sv, sa):
        return sa * sv / (sv + (sa - sv) * s) ** 2
    monkeypatch.setattr(N, "time_shift_sigma", shift_sigma)
    monkeypatch.setattr(N, "time_shift_slope", shift_slope)

    def net(x, sigma, cond):
        return 0.7 * x + 0.1 * torch.tanh(x) + cond
    net.sigma_shift_video, net.sigma_shift_audio = 3.0, 1.5
    g = torch.Generator().manual_seed(12)
    n_video, n_audio = 96, 32
    y = torch.randn(1, 8, n_video + n_audio, generator=g)
    noise = torch.randn(1, 8, n_video + n_audio, generator=g)
    noise_mask = (torch.rand(1, 8, n_video + n_audio, generator=g) < 0.5).float()
    sig = torch.tensor([0.95, 0.8, 0.6, 0.4, 0.2, 0.0])

    patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(net, model_type=minicomfy.ModelType.FLOW), dev)
    N._set_hyper(patcher, num_steps=3, cfg=1.0, prompt_mode="Image First")
    guider = minicomfy.CFGGuider(patcher)
    guider.set_conds(0.3, -0.2)
    guider.set_cfg(1.0)
    torch.manual_seed(4)
    with N.override_sample_function():
        got = type(guider).outer_sample(guider, noise, y, minicomfy.ksampler("euler"), sig, denoise_mask=noise_mask,
                                        seed=2, latent_shapes=[(1, n_video), (1, n_audio)])
    assert N.LAST_RUN["fused"] is False and torch.isfinite(got).all()

    # the same run against the engine seam: tensor mask, explicit audio arguments per outer step
    class Guider:
        inner_model = None

        def __init__(self):
            self.inner_model = self
            self.model_sampling = minicomfy.ModelSamplingCONST()

        def __call__(self, x, t, model_options=None, seed=None):
            return net(x, t, 0.3)
    eng = LanPaint(Guider(), 3, 15.0, 5.0, 1.0, 0.2, IS_FLOW=True, MinStepFrac=1.0)
    torch.manual_seed(4)
    yd, nd = y.to(dev), noise.to(dev)
    x = Guider().model_sampling.noise_scaling(sig[0].to(dev), nd, yd)
    ind = torch.zeros_like(x)
    ind[..., n_video:] = 1.0
    known = (1.0 - noise_mask).to(dev)
    host = [float(v) for v in sig]
    for i in range(len(host) - 1):
        s = torch.full((1,), host[i])
        times = times_from_sigma(s, True)
        n_eff = N.effective_inner_steps(3, host, host[i], float((1 - times[1]).mean()), 1, 1.0)
        fa = shift_sigma(times[2], 3.0, 1.5)
        abt_a = (1 - fa) ** 2 / ((1 - fa) ** 2 + fa ** 2)
        c = float(fa) / (host[i] * float(shift_slope(times[2], 3.0, 1.5))) if host[i] > 1e-4 else 1.0
        den = eng(x, yd, nd, s, known, times, {}, 2, n_steps=n_eff, current_times_audio=(fa / (1 - fa), abt_a, fa),
                  audio_indicator=ind, audio_correction=(1.0 - ind) + c * ind)
        x = x + (x - den) / host[i] * (host[i + 1] - host[i])
    want = Guider().model_sampling.inverse_noise_scaling(sig[-1].to(dev), x)
    assert max_rel(got, want) <= 1e-5


def test_a_real_torch_module_under_the_node_graphs(cuda_device):
    """The network is an nn.Module with cuDNN convolutions (what ComfyUI's diffusion_model is): it is captured inside
    the node path's graphs; in-place weight updates are picked up by replays (graphs read weights by address); moving
    the weights to new storage (ComfyUI off-loading / re-loading the model) drops the cached graphs."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    N._ENGINES.clear()
    dev = cuda_device

    class Net(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.body = torch.nn.Sequential(torch.nn.Conv2d(4, 8, 3, padding=1), torch.nn.SiLU(),
                                            torch.nn.Conv2d(8, 4, 3, padding=1))

        def forward(self, x, sigma, cond):
            return 0.6 * x + 0.1 * self.body(x) + cond
    torch.manual_seed(0)
    net = Net().to(dev)
    g = torch.Generator().manual_seed(4)
    y = torch.randn(2, 4, 32, 32, generator=g)
    noise_mask = (torch.rand(2, 1, 32, 32, generator=g) < 0.5).float()
    patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(net), dev)

    def run(p=patcher, opts=None):
        with torch.no_grad():
            out, _ = _ksampler_run(N, dev, y, noise_mask, patcher=p, opts=opts, steps=8, n=3)
        return out, N.LAST_RUN["mode"]

    outs, modes = zip(*[run() for _ in range(5)])
    assert modes[0] == "eager" and modes[-1] in ("steps", "job")
    for o in outs[1:]:
        assert max_rel(o, outs[0]) <= 1e-5        # cuDNN may pick another algorithm under capture: not bitwise
    # in-place weight update: same storage, the replayed graph must see the new values
    with torch.no_grad():
        for prm in net.parameters():
            prm.mul_(1.25)
    replayed, mode = run()
    assert mode == modes[-1]
    fresh, _ = run(minicomfy.ModelPatcher(minicomfy.BaseModel(net), dev), {"cuda_graph": False})
    assert max_rel(replayed, fresh) <= 1e-5 and max_rel(replayed, outs[0]) > 1e-3
    # new storage for the weights: graphs captured against the old addresses must not be replayed
    before = tuple(prm.data_ptr() for prm in net.parameters())
    net.to("cpu")
    hold = [torch.empty(4096, device=dev) for _ in range(64)]     # occupy the freed blocks: force new addresses
    net.to(dev)
    moved, mode = run()
    del hold
    if tuple(prm.data_ptr() for prm in net.parameters()) != before:
        assert mode == "eager"
    assert max_rel(moved, fresh) <= 1e-5


def test_whole_sampler_loop_as_one_graph(cuda_device):
    """heun (a deterministic sampler that is not plain Euler): the first job records what the sampler asks the wrapper
    for, the second captures k-diffusion's own loop -- its Python runs once -- into ONE graph, later jobs replay it.
    Same latent as plain launches bit for bit, same number of model calls, every progress callback delivered."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    N._ENGINES.clear()
    dev = cuda_device
    g = torch.Generator().manual_seed(5)
    y = torch.randn(2, 4, 32, 32, generator=g)
    noise_mask = (torch.rand(2, 1, 32, 32, generator=g) < 0.5).float()
    plain, _ = _ksampler_run(N, dev, y, noise_mask, opts={"cuda_graph": False}, sampler="heun", steps=10, n=3)
    calls = N.LAST_ENGINE["engine"].model_calls
    patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(_denoiser), dev)
    modes = []
    for k in range(4):
        before = minicomfy.PROGRESS["calls"]
        got, _ = _ksampler_run(N, dev, y, noise_mask, patcher=patcher, sampler="heun", steps=10, n=3)
        modes.append(N.LAST_RUN["mode"])
        assert torch.equal(got, plain), (k, modes)
        assert N.LAST_ENGINE["engine"].model_calls == calls
        assert minicomfy.PROGRESS["calls"] - before == 10 and minicomfy.PROGRESS["last"] == (10, 10)
    assert modes == [None, "sampler-graph", "sampler-graph", "sampler-graph"], modes
    assert N.LAST_RUN["job"].captures == 1
    # another seed: same graph, fresh noise image and randn stream
    other, _ = _ksampler_run(N, dev, y, noise_mask, patcher=patcher, sampler="heun", steps=10, n=3, seed=99)
    ref, _ = _ksampler_run(N, dev, y, noise_mask, opts={"cuda_graph": False}, sampler="heun", steps=10, n=3, seed=99)
    assert N.LAST_RUN["mode"] is None and torch.equal(other, ref) and not torch.equal(other, plain)


def test_a_sampler_that_draws_its_own_noise_is_not_captured(cuda_device):
    """euler_ancestral calls torch.randn_like between wrapper calls: its generator consumption interleaves with the
    engine's, so the loop is not captured as a whole; the per-call graphs serve it and the result equals plain launches."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    N._ENGINES.clear()
    dev = cuda_device
    g = torch.Generator().manual_seed(7)
    y = torch.randn(1, 4, 32, 32, generator=g)
    noise_mask = (torch.rand(1, 1, 32, 32, generator=g) < 0.5).float()
    plain, _ = _ksampler_run(N, dev, y, noise_mask, opts={"cuda_graph": False}, sampler="euler_ancestral", steps=8, n=3)
    patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(_denoiser), dev)
    for k in range(3):
        got, _ = _ksampler_run(N, dev, y, noise_mask, patcher=patcher, sampler="euler_ancestral", steps=8, n=3)
        assert N.LAST_RUN["mode"] is None and torch.equal(got, plain), k


def test_flux_shaped_transformer_under_the_node_graphs(cuda_device):
    """BASELINE configs[3] in miniature: a Flux-type model (cfg 1 -> uncond skipped, cfg_BIG = 1, flow `simple`
    schedule) whose network is a transformer (minicomfy.networks.DiTStandIn: patchify, adaLN blocks, SDPA attention)
    captured in the node path's graphs.  Eager first job, captured second, replayed third: each equals the oracle's
    restatement of the reference run around the same network with the same seed."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    from minicomfy.networks import DiTStandIn
    N._ENGINES.clear()
    dev = cuda_device
    torch.manual_seed(0)
    net = DiTStandIn(in_ch=16, hidden=128, depth=2, heads=4, dtype=torch.float32).to(dev).eval()
    g = torch.Generator().manual_seed(12)
    y = torch.randn(1, 16, 16, 16, generator=g)
    noise_mask = (torch.rand(1, 1, 16, 16, generator=g) < 0.5).float()
    base = minicomfy.BaseModel(net, model_type=minicomfy.ModelType.FLUX, latent_channels=16, shift=1.15)
    patcher = minicomfy.ModelPatcher(base, dev)
    outs, modes = [], []
    with torch.no_grad():
        for _ in range(3):
            (out,) = N.LanPaint_KSampler().sample(patcher, 21, 10, 1.0, "euler", "simple", 0.3, -0.2,
                                                  {"samples": y, "noise_mask": noise_mask}, 1.0, 3, "Image First", "",
                                                  N.IMAGE_MODE)
            outs.append(out["samples"])
            modes.append(N.LAST_RUN["mode"])
        eng = N.LAST_ENGINE["engine"]
        assert modes[0] == "eager" and modes[-1] in ("steps", "job"), modes
        assert float(patcher.LanPaint_cfg_BIG) == 1.0 or eng.IS_FLUX        # nodes.py:331-334: Flux runs with cfg_BIG = 1

        class Guider:       # cfg == 1: one conditional evaluation serves both heads (nodes.py:161-175)
            def __init__(self):
                self.inner_model, self.model_sampling, self.calls = self, O.FlowSampling(), 0

            def __call__(self, x, t, model_options=None, seed=None):
                self.calls += 1
                o = net(x, t, 0.3)
                return o, o
        noise = minicomfy.prepare_noise(y, 21)      # re-seeds every generator exactly like the node call did
        sig = minicomfy.KSampler(patcher, 10, dev, "euler", "simple").sigmas.to(dev)
        model = Guider()
        want = O.euler_inpaint(model, y.to(dev), noise.to(dev), noise_mask.expand(1, 16, 16, 16).to(dev), sig,
                               O.Hyper(n_steps=3, min_step_frac=1.0, flow=True))
        want = base.model_sampling.inverse_noise_scaling(sig[-1], want)
    assert eng.model_calls == model.calls
    for o in outs:
        assert max_rel(o, want) <= 1e-4, max_rel(o, want)


def test_result_reaches_the_host_through_pinned_memory(cuda_device):
    """ComfyUI finishes a sample call with `samples.to(intermediate_device())`; the patched outer_sample has already
    put the result in a pinned tensor of the caching host allocator (same bytes, a fraction of the time), so that
    `.to` is a no-op.  `{"pinned_result": False}` leaves the copy to ComfyUI; results below 64 KiB are left alone."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    dev = cuda_device
    g = torch.Generator().manual_seed(9)
    y = torch.randn(2, 4, 64, 64, generator=g)                  # 128 KiB
    noise_mask = (torch.rand(2, 1, 64, 64, generator=g) < 0.5).float()
    pinned, _ = _ksampler_run(N, dev, y, noise_mask, opts={"cuda_graph": False}, steps=6, n=2)
    plain, _ = _ksampler_run(N, dev, y, noise_mask, opts={"cuda_graph": False, "pinned_result": False}, steps=6, n=2)
    assert pinned.device.type == "cpu" and pinned.is_pinned() and not plain.is_pinned()
    assert torch.equal(pinned, plain)
    small, _ = _ksampler_run(N, dev, y[:1, :, :16, :16].contiguous(), noise_mask[:1, :, :16, :16].contiguous(),
                             opts={"cuda_graph": False}, steps=6, n=2)
    assert small.device.type == "cpu" and not small.is_pinned()
    # the helper itself: non-contiguous input, another dtype, a CPU tensor passes through
    t = torch.randn(64, 1024, device=dev).t()
    h = N._host_result(t)
    assert h.is_pinned() and torch.equal(h, t.cpu()) and N._host_result(h) is h
    hb = N._host_result(torch.ones(1 << 16, dtype=torch.bfloat16, device=dev))
    assert hb.dtype == torch.bfloat16 and hb.is_pinned() and float(hb.float().sum()) == float(1 << 16)
