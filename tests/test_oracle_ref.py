"""oracle/_ref (the reference's own engine, compiled to bytecode by oracle/build_ref.py) against the oracle port:
the thing bench.py times as `cpu_baseline.kind == "reference"` is the code the port restates, bit for bit."""
import pytest
import torch

from oracle import build_ref
from oracle import langevin_oracle as O


@pytest.mark.parametrize("flow,n,batch", [(False, 5, 1), (False, 3, 2), (True, 4, 1)])
def test_reference_bytecode_equals_the_port(flow, n, batch, monkeypatch):
    Ref = build_ref.load()
    if Ref is None:
        pytest.skip("oracle/_ref not built (needs /root/reference: `make -C oracle`)")
    g = torch.Generator().manual_seed(3)
    shape = (batch, 4, 16, 16)
    x, y, noise = (torch.randn(shape, generator=g) for _ in range(3))
    mask = (torch.rand((batch, 1, 16, 16), generator=g) < 0.5).float().expand(shape).contiguous()
    sigma = torch.full((batch,), 0.6 if flow else 2.5)
    times = O.times_from_sigma(sigma, flow)
    hp = O.Hyper(n_steps=n, min_step_frac=1.0, flow=flow)
    sampling = O.FlowSampling() if flow else O.VESampling()
    tape = O.NoiseTape(generator=torch.Generator().manual_seed(4))
    want_out, want_x = O.outer_step(O.PointwiseDenoiser(sampling), x.clone(), y, noise, sigma, mask, times, hp,
                                    n_steps=n, draw=tape)
    replay = iter(tape.recorded)
    monkeypatch.setattr(torch, "randn_like", lambda t, **kw: next(replay).to(t.dtype))
    eng = Ref(O.PointwiseDenoiser(sampling), n, 15.0, hp.lam, hp.beta, hp.step_size, IS_FLUX=False, IS_FLOW=flow,
              MinStepFrac=1.0)
    xr = x.clone()
    out = eng(xr, y, noise, sigma, mask, tuple(times), {}, 0, n_steps=n)
    assert torch.equal(out, want_out) and torch.equal(xr, want_x)
