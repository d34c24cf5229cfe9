"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import torch

from oracle import langevin_oracle as O


def make_model(kind: str, flow: bool):
    sampling = O.FlowSampling() if flow else O.VESampling()
    if kind == "identity":
        return O.IdentityDenoiser(sampling)
    if kind == "two_heads":
        return O.PointwiseDenoiser(sampling)
    if kind == "bare":
        return O.PointwiseDenoiser(sampling, heads=0)
    if kind == "one_tuple":
        return O.PointwiseDenoiser(sampling, heads=1)
    raise KeyError(kind)


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.double().cpu()
    b = b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_rel(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max|b| : the 'relative fp32 on the final latent' of BASELINE.json."""
    a = a.double().cpu()
    b = b.double().cpu()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def synth_inputs(shape, seed=0, density=0.5, device="cpu", channel_mask=True):
    """SURVEY 8d synthetic inputs: x, y, noise ~ N(0,1); mask = rand(B,1,*sp) < density (1 = known)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(shape, generator=g)
    y = torch.randn(shape, generator=g)
    noise = torch.randn(shape, generator=g)
    mshape = (shape[0], 1) + tuple(shape[2:])
    if density >= 1:
        m = torch.ones(mshape)
    elif density <= 0:
        m = torch.zeros(mshape)
    else:
        m = (torch.rand(mshape, generator=g) < density).float()
    if not channel_mask:
        m = m.expand(shape).contiguous()
    return [t.to(device) for t in (x, y, noise, m)]
