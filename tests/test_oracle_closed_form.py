"""Two independently written CPU oracles must agree: oracle/closed_form.c (plain C, scalar, the per-element
closed form of SURVEY 8a that the CUDA kernels implement) against oracle/langevin_oracle.py (tensor ops in the
reference's order, pinned bit-for-bit to the reference).  fp64, so agreement is to round-off.  CPU only."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle import langevin_oracle as O


class Params(C.Structure):
    _fields_ = [("step_size", C.c_double), ("lam", C.c_double), ("beta", C.c_double), ("min_step_frac", C.c_double),
                ("flow", C.c_int32), ("n_steps", C.c_int32), ("coef", C.c_double * 5)]


@pytest.fixture(scope="module")
def clib():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_build", "libclosed_form.so"))
    lib.cf_outer_step.restype = C.c_int
    return lib


@pytest.mark.parametrize("flow,B,n,lam,beta,step,min_frac", [
    (False, 1, 5, 5.0, 1.0, 0.2, 1.0), (False, 3, 3, 8.0, 1.5, 0.15, 0.4), (True, 2, 4, 5.0, 1.0, 0.2, 1.0),
    (False, 1, 0, 5.0, 1.0, 0.2, 1.0), (True, 1, 1, 0.1, 0.5, 0.6, 0.0), (False, 2, 10, 5.0, 1.0, 0.2, 1.0)])
def test_c_closed_form_equals_tensor_oracle(clib, flow, B, n, lam, beta, step, min_frac):
    g = torch.Generator().manual_seed(B * 100 + n)
    shape = (B, 3, 5, 4)
    per = 60
    x, y, noise = (torch.randn(shape, generator=g, dtype=torch.float64) for _ in range(3))
    mask = (torch.rand(shape, generator=g) < 0.5).double()          # element-wise, channel-dependent
    sig = torch.rand(B, generator=g, dtype=torch.float64) * (0.9 if flow else 6.0) + 0.05
    if B == 1:
        sig = sig[:1]
    times = O.times_from_sigma(sig, flow)
    coef = (0.7, 0.1, 0.0, 0.6, -0.05)
    model = O.PointwiseDenoiser(O.FlowSampling() if flow else O.VESampling(), coef=coef)
    hp = O.Hyper(n_steps=n, lam=lam, beta=beta, step_size=step, min_step_frac=min_frac, flow=flow)
    tape = O.NoiseTape(generator=torch.Generator().manual_seed(7))
    tape_draws = [torch.randn(shape, generator=torch.Generator().manual_seed(50 + k), dtype=torch.float64)
                  for k in range(max(1, 2 * n - 1))]
    out_t, x_t = O.outer_step(model, x.clone(), y, noise, sig, mask, times, hp, n_steps=n, draw=O.NoiseTape(tape_draws))
    # replace form exactly as the reference picks it (lanpaint.py:85-92): noise_scaling for one sigma, flow form else
    if sig.numel() == 1:
        rn, ry = ([float(sig[0])], [1.0 - float(sig[0])]) if flow else ([float(sig[0])], [1.0])
    else:
        rn, ry = [float(s) for s in sig], [1.0 - float(s) for s in sig]
    p = Params(step, lam, beta, min_frac, int(flow), n, (C.c_double * 5)(*coef))
    arr = lambda t: np.ascontiguousarray(t.reshape(B, per).numpy())
    xa, ya, na = arr(x), arr(y), arr(noise)
    ma = np.ascontiguousarray(mask.reshape(B, per).numpy().astype(np.uint8))
    tp = np.ascontiguousarray(np.stack([arr(d) for d in tape_draws]))
    ab, vv = (np.ascontiguousarray(t.numpy().astype(np.float64)) for t in (times.abt, times.ve_sigma))
    rn_a, ry_a = np.asarray(rn, np.float64), np.asarray(ry, np.float64)
    out_c, x_c = np.empty_like(xa), np.empty_like(xa)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    used = clib.cf_outer_step(P(xa), P(ya), P(na), P(ma), P(ab), P(vv), P(rn_a), P(ry_a), P(tp), C.c_int64(B),
                              C.c_int64(per), C.byref(p), P(out_c), P(x_c))
    assert used == (0 if n == 0 else 2 * n - 1)
    scale = float(x_t.abs().max())
    assert np.abs(x_c - arr(x_t)).max() <= 1e-11 * scale
    assert np.abs(out_c - arr(out_t)).max() <= 1e-11 * max(scale, float(out_t.abs().max()))
