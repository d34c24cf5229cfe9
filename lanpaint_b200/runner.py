"""Host-scheduled sampling runs over the engine.

`HostSchedule` does on the host, once per run, what the reference's per-sigma
wrapper recomputes on the device every outer step with two syncs
(src/LanPaint/nodes.py:242-252,286-299): the (VE sigma, abt, flow t) triple and
the effective inner-step count for every sigma of the schedule.
`euler_inpaint` is then the reference's `LanPaint_KSampler` run with sampler
"euler" (nodes.py:338,376-378 around k-diffusion's sample_euler) expressed as a
sync-free launch sequence.

`SynthDenoiser` is the SURVEY 8d pointwise two-head stand-in for the UNet/DiT,
running as one CUDA kernel (lp_synth_denoiser_f32); bench.py and the tests use it.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from . import _native
from .schedule import effective_inner_steps, times_from_sigma


class VESampling:
    """EPS-type model_sampling: noise_scaling = y + sigma*noise (what nodes.py:338 relies on)."""

    def noise_scaling(self, sigma, noise, latent_image, max_denoise=False):
        if max_denoise:
            return noise * torch.sqrt(1.0 + sigma ** 2.0) + latent_image
        return latent_image + noise * sigma

    def inverse_noise_scaling(self, sigma, latent):
        return latent


class FlowSampling:
    """CONST-type (rectified flow) model_sampling."""

    noise_scale = 1.0

    def noise_scaling(self, sigma, noise, latent_image, max_denoise=False):
        return sigma * (self.noise_scale * noise) + (1.0 - sigma) * latent_image

    def inverse_noise_scaling(self, sigma, latent):
        return latent / (1.0 - sigma)


class SynthDenoiser:
    """x -> (a0 x + b0 tanh x + c0, a1 x + c1) as ONE kernel writing both heads.

    Follows the model protocol of the engine seam (`inner_model.model_sampling`,
    `__call__(x, t, model_options=, seed=)`).  Output buffers are reused, which is
    what a CUDA-graph-captured network does too."""

    def __init__(self, sampling=None, coef=(0.7, 0.1, 0.0, 0.6, -0.05), two_heads: bool = True):
        self.inner_model = self
        self.model_sampling = sampling or VESampling()
        self.coef = tuple(float(c) for c in coef)
        self._coef_c = (C.c_float * 5)(*self.coef)
        self.two_heads = two_heads
        self.calls = 0
        self._h0: Optional[torch.Tensor] = None
        self._h1: Optional[torch.Tensor] = None
        self._lib = _native.load()

    def set_coef(self, coef: Sequence[float]):
        self.coef = tuple(float(c) for c in coef)
        self._coef_c = (C.c_float * 5)(*self.coef)

    def __call__(self, x, t, model_options=None, seed=None):
        if self._h0 is None or self._h0.shape != x.shape or self._h0.device != x.device:
            self._h0 = torch.empty_like(x)
            self._h1 = torch.empty_like(x) if self.two_heads else None
        rc = self._lib.lp_synth_denoiser_f32(C.c_void_p(x.data_ptr()), C.c_void_p(self._h0.data_ptr()),
                                             C.c_void_p(self._h1.data_ptr()) if self._h1 is not None else None,
                                             x.numel(), self._coef_c,
                                             C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
        _native.check(rc, "lp_synth_denoiser_f32")
        self.calls += 1
        return (self._h0, self._h1) if self.two_heads else self._h0


@dataclass
class OuterStep:
    sigma: float
    sigma_next: float
    n_inner: int
    sigma_t: torch.Tensor    # CPU [B] fp32, what the sampler would pass as `sigma`
    times: tuple             # CPU (VE_Sigma, abt, Flow_t), reference op order (nodes.py:242-252)


class HostSchedule:
    """All per-sigma scalars of a run, computed once on the host."""

    def __init__(self, sigmas: Sequence[float], batch: int, n_inner: int, flow: bool = False,
                 early_stop: int = 1, min_step_frac: float = 1.0):
        sig = torch.as_tensor(list(sigmas), dtype=torch.float32)
        self.sigmas = sig
        host = [float(v) for v in sig]
        self.steps: List[OuterStep] = []
        ones = torch.ones(batch, dtype=torch.float32)
        for i in range(len(host) - 1):
            s = sig[i] * ones
            tm = times_from_sigma(s, flow)
            n_eff = effective_inner_steps(n_inner, host, float(torch.mean(s)), float((1.0 - tm[1]).mean()),
                                          early_stop, min_step_frac)
            self.steps.append(OuterStep(host[i], host[i + 1], n_eff, s, tm))
        self.substeps = sum(st.n_inner for st in self.steps)
        self.model_calls = self.substeps + len(self.steps)


def karras_sigmas(n: int, sigma_min: float = 0.0292, sigma_max: float = 14.6146, rho: float = 7.0) -> List[float]:
    """SDXL's karras schedule, trailing 0 included."""
    ramp = torch.linspace(0, 1, n, dtype=torch.float32)
    lo, hi = sigma_min ** (1 / rho), sigma_max ** (1 / rho)
    s = (hi + ramp * (lo - hi)) ** rho
    return [float(v) for v in s] + [0.0]


def euler_inpaint(engine, latent_image: torch.Tensor, noise: torch.Tensor, mask, sched: HostSchedule,
                  x_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One full inpaint job on device-resident inputs; returns the final latent.

    mask: latent_mask (1 = known) tensor or an engine PackedMask."""
    sampling = engine.inner_model.inner_model.model_sampling
    x = sampling.noise_scaling(sched.sigmas[0].to(latent_image.device), noise, latent_image)
    for st in sched.steps:
        denoised = engine(x, latent_image, noise, st.sigma_t, mask, st.times, None, 0, n_steps=st.n_inner)
        # k-diffusion sample_euler: d = (x - denoised)/sigma ; x = x + d*(sigma_next - sigma)
        x = torch.add(x, x - denoised, alpha=(st.sigma_next - st.sigma) / st.sigma)
    if x_out is not None:
        x_out.copy_(x, non_blocking=True)
        return x_out
    return x
