"""Property tests (hypothesis) of the host coefficient table: identities the exact OU solution must satisfy for
every admissible (abt, sigma, hyper-parameters), independent of any oracle.  CPU only."""
import math

import numpy as np
from hypothesis import given, settings, strategies as st

from lanpaint_b200 import schedule

F = st.floats


@settings(max_examples=300, deadline=None)
@given(abt=F(1e-4, 1 - 1e-4), sigma=F(0.01, 50.0), lam=F(0.1, 50.0), beta=F(0.1, 4.0), step=F(1e-4, 1.0),
       min_frac=F(0.0, 1.0), flow=st.booleans())
def test_ou_identities(abt, sigma, lam, beta, step, min_frac, flow):
    abt32 = float(np.float32(abt))
    hp = schedule.Hyper(step, lam, beta, min_frac, flow)
    t = schedule.build_table([abt32], [sigma], hp, [0.3], [0.7], [1.0])[0].astype(np.float64)
    one_m = 1.0 - abt32
    h = step * max(one_m, min_frac)
    rel = lambda a, b: abs(a - b) <= 3e-6 * max(abs(a), abs(b), 1e-30) + 1e-30
    assert rel(t[1] * t[2], 1.0)                                   # S * inv_S
    assert rel(t[0], math.sqrt(abt32) / one_m)                     # c_tgt
    assert rel(t[4], 1.0 + t[3]) and rel(t[3], lam)
    for k, (A, dt) in enumerate(((1.0 / one_m, h), ((1.0 + lam) / one_m, h * beta))):
        c = t[8 + 8 * k: 16 + 8 * k]
        g, dtt, ef, kf, sf, eh, kh, sh = c
        assert rel(g + 1.0 / one_m, A) and rel(dtt, dt)
        for e, kk, sd, hh in ((ef, kf, sf, dt), (eh, kh, sh, dt / 2)):
            assert 0.0 <= e <= 1.0 and kk >= 0.0 and sd >= 0.0
            assert abs(e + A * kk - 1.0) <= 1e-5                   # k = (1 - e)/A
            assert abs(sd * sd - (1.0 - math.exp(-2 * A * hh)) / A) <= 1e-5 * max(1.0, 1.0 / A)   # D^2 = 2
            assert rel(e, math.exp(-A * hh)) or e < 1e-30
        assert abs(eh * eh - ef) <= 1e-6                           # two half steps decay like one full step
        sdm, sdmf = t[24 + k], t[26 + k]
        assert abs(sdm * sdm - ((eh * sh) ** 2 + sh * sh)) <= 1e-5 * max(1.0, sdm * sdm)
        assert abs(sdmf * sdmf - ((eh * sf) ** 2 + sh * sh)) <= 1e-5 * max(1.0, sdmf * sdmf)
        # variance bookkeeping of the exact solution: a full step == two half steps
        assert abs(sf * sf - (sh * sh * (1 + eh * eh))) <= 1e-5 * max(1.0, sf * sf)


@settings(max_examples=200, deadline=None)
@given(n=st.integers(0, 100), frac=F(0.0, 1.0), min_frac=F(0.0, 1.0))
def test_inner_step_ramp_properties(n, frac, min_frac):
    got = schedule.min_step_frac_effective_steps(n, frac, min_frac)
    assert 0 <= got <= n
    if min_frac <= 0 or frac >= min_frac:
        assert got == n
    else:
        assert got == max(0, round(n * frac / min_frac))
