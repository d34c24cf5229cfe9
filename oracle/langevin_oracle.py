"""CPU oracle for LanPaint's inner Langevin "think" loop.

THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl
reference`` legs may import it.  The product path (``lanpaint_b200``) never
does; it fails loudly when the CUDA library is missing.

What it is: a from-scratch restatement, in plain tensor arithmetic, of the
algorithm in the reference's ``src/LanPaint/lanpaint.py`` (the engine) and of
the schedule glue in ``src/LanPaint/nodes.py`` that feeds it.  Every function
cites the reference lines it restates.  It keeps the reference's *operation
order* (one full-tensor op per arithmetic step, scalars broadcast as
``[B,1,1,..]`` tensors) so that

* in fp32 it reproduces the reference bit-for-bit on the same noise tape
  (pinned by ``tests/test_oracle_golden.py`` against fixtures generated from
  the real reference by ``tests/golden/make_golden.py``), and
* timing it on the host cores is a fair stand-in for the reference's own
  eager-PyTorch CPU path (same ~90 full-tensor passes per sub-step).

It is dtype- and device-agnostic (fp32/fp64; ``cpu`` or ``cuda``), draws noise
through an injectable ``draw(like) -> tensor`` (default ``torch.randn_like``,
i.e. the global generator exactly like ``lanpaint.py:252``), and never
mutates its inputs: ``outer_step`` returns ``(out, x_new)`` where the
reference returns ``out`` and writes ``x_new`` into ``x`` in place
(``lanpaint.py:156``).

Parity status: PINNED (golden vectors from the reference itself + the
reference tests' known answers, see tests/test_oracle_golden.py).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, NamedTuple, Optional, Sequence

import torch

Tensor = torch.Tensor
Draw = Callable[[Tensor], Tensor]


# --------------------------------------------------------------------------
# small containers
# --------------------------------------------------------------------------
@dataclass
class Hyper:
    """Constructor arguments of the reference engine (lanpaint.py:8-21)."""

    n_steps: int = 5
    friction: float = 15.0  # computed into Gamma_* but unused by the live scheme
    lam: float = 5.0
    beta: float = 1.0
    step_size: float = 0.2
    min_step_frac: float = 0.0  # engine default; the nodes pass 1.0
    flow: bool = False  # IS_FLUX or IS_FLOW


class Times(NamedTuple):
    """``current_times`` triple of lanpaint.py:58 / nodes.py:283."""

    ve_sigma: Tensor
    abt: Tensor
    flow_t: Tensor


class State(NamedTuple):
    """LangevinState (types.py:6-9); ``v`` is always None."""

    v: Optional[Tensor]
    C: Optional[Tensor]
    x0: Optional[Tensor]


@dataclass
class Audio:
    """MiniMax-H3 per-row schedule context (lanpaint.py:60-74, 173-180)."""

    indicator: Tensor  # 1 on audio rows, broadcastable to the latent
    times: Times  # audio-stream (VE, abt, flow_t)
    correction: Optional[Tensor] = None  # c per row, or None


# --------------------------------------------------------------------------
# schedule glue (nodes.py)
# --------------------------------------------------------------------------
def times_from_sigma(sigma: Tensor, flow: bool) -> Times:
    """sigma -> (VE sigma, alpha-bar, flow t).  nodes.py:242-252."""
    if flow:
        t = sigma
        abt = (1 - t) ** 2 / ((1 - t) ** 2 + t ** 2)
        ve = t / (1 - t)
        return Times(ve, abt, t)
    ve = sigma
    abt = 1 / (1 + ve ** 2)
    t = (1 - abt) ** 0.5 / ((1 - abt) ** 0.5 + abt ** 0.5)
    return Times(ve, abt, t)


def ramped_inner_steps(n_steps: int, frac: float, min_frac: float) -> int:
    """MinStepFrac tail ramp, Python banker's rounding.  nodes.py:134-144."""
    if min_frac <= 0 or frac >= min_frac or n_steps <= 0:
        return n_steps
    return max(0, round(n_steps * frac / min_frac))


def inner_steps_for(sigma: Tensor, sigmas: Tensor, abt: Tensor, n_steps: int,
                    early_stop: int = 1, min_frac: float = 1.0) -> int:
    """How many sub-steps the per-sigma wrapper requests.  nodes.py:286-299."""
    current = int(torch.argmin(torch.abs(sigmas - torch.mean(sigma))))
    total = len(sigmas) - 1
    if total - current <= early_stop:
        return 0
    return ramped_inner_steps(n_steps, float((1.0 - abt).mean()), min_frac)


def binarise_mask(denoise_mask: Tensor) -> Tensor:
    """latent_mask = 1 - (denoise_mask > 0.5).  nodes.py:281-283."""
    return 1 - (denoise_mask > 0.5).float()


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
def _lift(t: Tensor, ndim: int) -> Tensor:
    """add_none_dims (lanpaint.py:23-29): append singleton dims up to ndim."""
    while t.ndim < ndim:
        t = t.unsqueeze(t.ndim)
    return t


def _first_col(t: Tensor, ndim: int) -> Tensor:
    """remove_none_dims (lanpaint.py:30-33)."""
    return t[(slice(None),) + (0,) * (ndim - 1)]


def split_heads(output):
    """unpack_model_output (lanpaint.py:34-43)."""
    if isinstance(output, (tuple, list)):
        if len(output) >= 2:
            return output[0], output[1]
        if len(output) == 1:
            return output[0], output[0]
        raise ValueError("Model output is empty")
    return output, output


# --------------------------------------------------------------------------
# coefficients (lanpaint.py:295-328)
# --------------------------------------------------------------------------
class Coefs(NamedTuple):
    half_dt_x: Tensor
    half_dt_y: Tensor
    A_x: Tensor
    A_y: Tensor
    D_x: Tensor
    D_y: Tensor


def branch_coefficients(abt: Tensor, step: Tensor, sig_x: Tensor, sig_y: Tensor,
                        lam: float) -> Coefs:
    """Per-branch time step, stiffness and diffusion.  lanpaint.py:295-328.

    Keeps the reference's round trip A = (A*dt/2)/(dt/2) so fp32 bits match.
    """
    dtx = 2 * step * sig_x
    dty = 2 * step * sig_y
    Atx = (1) / (1 - abt) * dtx / 2
    Aty = (1 + lam) / (1 - abt) * dty / 2
    A_x = Atx / (dtx / 2)
    A_y = Aty / (dty / 2)
    D_x = (2 * abt ** 0) ** 0.5
    D_y = (2 * abt ** 0) ** 0.5
    return Coefs(dtx / 2, dty / 2, A_x, A_y, D_x, D_y)


# --------------------------------------------------------------------------
# one exact Ornstein-Uhlenbeck advance (lanpaint.py:232-254)
# --------------------------------------------------------------------------
def ou_advance(x: Tensor, h: Tensor, A: Tensor, C: Tensor, D: Tensor, draw: Draw) -> Tensor:
    Ah = A * h
    decay = torch.exp(-Ah)
    tiny = torch.abs(A) < 1e-8
    k = torch.where(tiny, h, (-torch.expm1(-Ah)) / A)
    k2 = torch.where(tiny, h, (-torch.expm1(-2 * Ah)) / (2 * A))
    mean = decay * x + k * C
    var = (D ** 2) * k2
    kick = draw(x) * torch.sqrt(torch.clamp(var, min=0.0))
    return mean + kick


# --------------------------------------------------------------------------
# masked score (lanpaint.py:159-184)
# --------------------------------------------------------------------------
def masked_score(model, x_t: Tensor, y: Tensor, mask: Tensor, abt: Tensor, ve: Tensor,
                 flow_t: Tensor, hp: Hyper, ndim: int, model_options=None, seed=None,
                 correction: Optional[Tensor] = None) -> Tensor:
    if hp.flow:
        x = x_t / (abt ** 0.5 + (1 - abt) ** 0.5)
        heads = model(x, _first_col(flow_t, ndim), model_options=model_options, seed=seed)
    else:
        x = x_t * (1 + ve ** 2) ** 0.5
        heads = model(x, _first_col(ve, ndim), model_options=model_options, seed=seed)
    x0, x0_big = split_heads(heads)
    if correction is not None:
        x0 = x + correction * (x0 - x)
        x0_big = x + correction * (x0_big - x)
    s_free = -(x_t - x0)
    s_known = -(1 + hp.lam) * (x_t - y) + hp.lam * (x_t - x0_big)
    return s_free * (1 - mask) + s_known * mask


# --------------------------------------------------------------------------
# one Langevin sub-step (lanpaint.py:192-293)
# --------------------------------------------------------------------------
def langevin_substep(x_t: Tensor, score: Callable[[Tensor], Tensor], mask: Tensor, step: Tensor,
                     times: Times, hp: Hyper, state: Optional[State], draw: Draw,
                     sig_x: Tensor, sig_y: Tensor):
    ndim = x_t.ndim
    abt = _lift(times.abt, ndim)
    cf = branch_coefficients(abt, step, sig_x, sig_y, hp.lam)
    if torch.mean(cf.half_dt_x) <= 0.0:
        return x_t, state
    A = cf.A_x * (1 - mask) + cf.A_y * mask
    D = cf.D_x * (1 - mask) + cf.D_y * mask
    dt = cf.half_dt_x * (1 - mask) + cf.half_dt_y * mask

    def drift_const(z: Tensor):
        x0e = z + score(z)
        C = (abt ** 0.5 * x0e - z) / (1 - abt) + A * z
        return C, x0e

    if state is None:
        C, x0e = drift_const(x_t)
        x_t = ou_advance(x_t, dt, A, C, D, draw)
    else:
        C = state.C
        x_t = ou_advance(x_t, dt / 2, A, C, D, draw)
        C_new, x0e = drift_const(x_t)
        x_t = x_t + (C_new - C) * dt
        x_t = ou_advance(x_t, dt / 2, A, C, D, draw)  # old C on purpose (lanpaint.py:283-284)
        C = C_new
    return x_t, State(None, C, x0e)


# --------------------------------------------------------------------------
# per-row schedule of one outer step (lanpaint.py:58-82)
# --------------------------------------------------------------------------
def per_row_schedule(sigma: Tensor, times: Times, hp: Hyper, audio: Optional[Audio], ndim: int):
    """(VE, abt, flow_t, replace sigma, step size [lifted], times, correction).  With an AV pack the audio
    positions take the audio stream's VE / abt and replace sigma (lanpaint.py:66-74); the step size is
    StepSize * clamp(1 - abt, min=MinStepFrac) per position (lanpaint.py:81-82)."""
    ve, abt, flow_t = times
    rep_sigma = sigma
    correction = None
    if audio is not None:
        ai = audio.indicator
        ve = ve * (1 - ai) + audio.times.ve_sigma * ai
        abt = abt * (1 - ai) + audio.times.abt * ai
        rep_sigma = sigma * (1 - ai) + audio.times.flow_t * ai
        times = Times(ve, abt, flow_t)
        correction = audio.correction
    step = hp.step_size * (1 - abt).clamp(min=hp.min_step_frac)
    step = _lift(step, ndim)
    return ve, abt, flow_t, rep_sigma, step, times, correction


# --------------------------------------------------------------------------
# one outer diffusion step (lanpaint.py:44-157)
# --------------------------------------------------------------------------
def outer_step(model, x: Tensor, y: Tensor, noise: Tensor, sigma: Tensor, mask: Tensor,
               times: Times, hp: Hyper, n_steps: Optional[int] = None, draw: Draw = torch.randn_like,
               model_options=None, seed=None, audio: Optional[Audio] = None,
               on_substep: Optional[Callable] = None):
    """Returns ``(out, x_new)``; the reference returns out and does x.copy_(x_new).

    ``on_substep(i, x_t, state) -> bool`` is a hook for the early-stop oracle
    (return True to break, lanpaint.py:122-142).
    """
    ndim = x.ndim
    if torch.mean(torch.abs(noise)) < 1e-8:  # lanpaint.py:51-52
        noise = draw(noise)
    if n_steps is None:
        n_steps = hp.n_steps

    ve, abt, flow_t, rep_sigma, step, times, correction = per_row_schedule(sigma, times, hp, audio, ndim)

    sampling = model.inner_model.model_sampling
    s = _lift(rep_sigma, ndim)
    if s.numel() == 1:  # lanpaint.py:85-92
        known = sampling.noise_scaling(s, noise, y)
    else:
        ns = getattr(sampling, "noise_scale", 1.0)
        known = s * (ns * noise) + (1.0 - s) * y
    x = x * (1 - mask) + known * mask

    abt_b = _lift(abt, ndim)
    ve_b = _lift(ve, ndim)
    if hp.flow:  # lanpaint.py:96-99
        x_t = x * (abt_b ** 0.5 + (1 - abt_b) ** 0.5)
    else:
        x_t = x / (1 + ve_b ** 2) ** 0.5

    sig_x = _lift(abt ** 0, ndim)  # lanpaint.py:185-190
    sig_y = _lift(hp.beta * abt ** 0, ndim)
    flow_b = _lift(flow_t, ndim)

    state: Optional[State] = None
    for i in range(n_steps):  # lanpaint.py:113-142
        def score(z, _abt=abt_b, _ve=ve_b, _ft=flow_b):
            return masked_score(model, z, y, mask, _abt, _ve, _ft, hp, ndim,
                                model_options=model_options, seed=seed, correction=correction)

        x_t, state = langevin_substep(x_t, score, mask, step, times, hp, state, draw, sig_x, sig_y)
        if on_substep is not None and on_substep(i, x_t, state):
            break

    if hp.flow:  # lanpaint.py:144-147
        x = x_t / (abt_b ** 0.5 + (1 - abt_b) ** 0.5)
    else:
        x = x_t * (1 + ve_b ** 2) ** 0.5

    out, _ = split_heads(model(x, sigma, model_options=model_options, seed=seed))  # lanpaint.py:151
    out = out * (1 - mask) + y * mask
    return out, x


# --------------------------------------------------------------------------
# a whole sampling run: Euler outer loop + per-sigma wrapper (nodes.py:229-300,
# 338, 376-378 with k-diffusion's sample_euler as the sampler_function)
# --------------------------------------------------------------------------
def euler_inpaint(model, latent_image: Tensor, noise: Tensor, denoise_mask: Tensor, sigmas: Tensor,
                  hp: Hyper, draw: Draw = torch.randn_like, early_stop: int = 1,
                  min_frac: float = 1.0, max_denoise: bool = False, counters: Optional[dict] = None):
    """Reference semantics of one `LanPaint_KSampler` run with sampler "euler".

    x starts as noise_scaling(sigmas[0], noise, latent_image) (nodes.py:338), each
    outer step calls the Langevin wrapper (which rewrites x in place) and then
    takes the Euler step from the rewritten x; inverse_noise_scaling at the end
    (nodes.py:378) is the identity for VE models and is applied by the caller
    for flow models.
    """
    sampling = model.inner_model.model_sampling
    if max_denoise:
        x = sampling.noise_scaling(sigmas[0], noise, latent_image, True)
    else:
        x = sampling.noise_scaling(sigmas[0], noise, latent_image)
    mask = binarise_mask(denoise_mask)
    B = x.shape[0]
    s_in = x.new_ones([B])
    sub = 0
    calls = 0
    for i in range(len(sigmas) - 1):
        sigma = sigmas[i] * s_in
        tm = times_from_sigma(sigma, hp.flow)
        n_eff = inner_steps_for(sigma, sigmas, tm.abt, hp.n_steps, early_stop, min_frac)
        denoised, x = outer_step(model, x, latent_image, noise, sigma, mask, tm, hp, n_eff, draw)
        sub += n_eff
        calls += n_eff + 1
        d = (x - denoised) / _lift(sigma, x.ndim)
        x = x + d * (sigmas[i + 1] - sigmas[i])
    if counters is not None:
        counters["substeps"] = sub
        counters["model_calls"] = calls
    return x


# --------------------------------------------------------------------------
# noise tape: make the RNG an explicit input so CPU and GPU can be compared
# --------------------------------------------------------------------------
class NoiseTape:
    """Replays a fixed list of draws (or records fresh ones).

    Draw order per outer step: [regen draw if noise~0],
    then 1 draw for sub-step 0 and 2 draws (first half, second half) for each
    later sub-step.
    """

    def __init__(self, draws: Optional[Sequence[Tensor]] = None, generator: Optional[torch.Generator] = None):
        self.draws = list(draws) if draws is not None else None
        self.recorded: list[Tensor] = []
        self.generator = generator
        self.pos = 0

    def __call__(self, like: Tensor) -> Tensor:
        if self.draws is not None:
            t = self.draws[self.pos].to(dtype=like.dtype, device=like.device)
            assert t.shape == like.shape, (t.shape, like.shape)
        else:
            t = torch.randn(like.shape, generator=self.generator, dtype=torch.float32).to(
                dtype=like.dtype, device=like.device)
        self.recorded.append(t)
        self.pos += 1
        return t


# --------------------------------------------------------------------------
# stand-in denoisers with the model protocol the engine expects
# (tests/test_av_schedule.py:110-130, tests/test_lanpaint_semantic_stop.py:6-17)
# --------------------------------------------------------------------------
class VESampling:
    """EPS-style noise_scaling: y + sigma * noise (nodes.py:338 comment)."""

    def noise_scaling(self, sigma, noise, latent_image, max_denoise=False):
        if max_denoise:
            return noise * torch.sqrt(1.0 + sigma ** 2.0) + latent_image
        return latent_image + noise * sigma

    def inverse_noise_scaling(self, sigma, latent):
        return latent


class FlowSampling:
    """CONST-style noise_scaling: sigma*noise + (1-sigma)*y."""

    noise_scale = 1.0

    def noise_scaling(self, sigma, noise, latent_image, max_denoise=False):
        return sigma * (self.noise_scale * noise) + (1.0 - sigma) * latent_image

    def inverse_noise_scaling(self, sigma, latent):
        return latent / (1.0 - sigma)


class PointwiseDenoiser:
    """x -> (a0*x + b0*tanh(x) + c0, a1*x + c1): two distinct heads, no spatial mixing."""

    def __init__(self, sampling, coef=(0.7, 0.1, 0.0, 0.6, -0.05), heads: int = 2):
        self.inner_model = self
        self.model_sampling = sampling
        self.coef = coef
        self.heads = heads
        self.calls = 0
        self.last_input = None

    def __call__(self, x, sigma, model_options=None, seed=None):
        self.calls += 1
        self.last_input = x
        a0, b0, c0, a1, c1 = self.coef
        h0 = a0 * x + b0 * torch.tanh(x) + c0
        if self.heads == 0:
            return h0  # bare tensor: both heads alias
        if self.heads == 1:
            return (h0,)
        return h0, a1 * x + c1


class IdentityDenoiser(PointwiseDenoiser):
    """BASELINE config 1's "dummy eps=identity denoiser": returns (x, x)."""

    def __init__(self, sampling):
        super().__init__(sampling, coef=(1.0, 0.0, 0.0, 1.0, 0.0))

    def __call__(self, x, sigma, model_options=None, seed=None):
        self.calls += 1
        self.last_input = x
        return x, x


def karras_sigmas(n: int, sigma_min: float = 0.0292, sigma_max: float = 14.6146, rho: float = 7.0) -> Tensor:
    """k-diffusion get_sigmas_karras (SDXL karras-20), with the trailing 0."""
    ramp = torch.linspace(0, 1, n)
    lo = sigma_min ** (1 / rho)
    hi = sigma_max ** (1 / rho)
    s = (hi + ramp * (lo - hi)) ** rho
    return torch.cat([s, s.new_zeros([1])])


def flow_simple_sigmas(n: int, shift: float = 1.0) -> Tensor:
    """A 'simple'-style rectified-flow schedule with time shift, trailing 0."""
    t = torch.linspace(1.0, 1.0 / n, n)
    s = shift * t / (1 + (shift - 1) * t)
    return torch.cat([s, s.new_zeros([1])])
