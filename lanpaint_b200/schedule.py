"""Host-side schedule logic of the Langevin hot path.

Everything the reference computes per outer step as device scalars-with-syncs
(src/LanPaint/nodes.py:242-252,286-299; src/LanPaint/lanpaint.py:81,205,
295-328) happens here once, on the host, from the per-sample sigma values:
the (VE sigma, alpha-bar, flow t) triple, the effective inner-step count and
the coefficient table the kernels read.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np
import torch

from . import _native

TABLE_STRIDE = _native.TABLE_STRIDE


def times_from_sigma(sigma: torch.Tensor, flow: bool):
    """sigma -> (VE_Sigma, abt, Flow_t) with the reference's own fp32 op order
    (nodes.py:242-252), so the alpha-bar the table is built from is bit-identical."""
    if flow:
        flow_t = sigma
        abt = (1 - flow_t) ** 2 / ((1 - flow_t) ** 2 + flow_t ** 2)
        ve = flow_t / (1 - flow_t)
    else:
        ve = sigma
        abt = 1 / (1 + ve ** 2)
        flow_t = (1 - abt) ** 0.5 / ((1 - abt) ** 0.5 + abt ** 0.5)
    return ve, abt, flow_t


def min_step_frac_effective_steps(n_steps, frac, min_frac):
    """Inner-step count under the MinStepFrac tail ramp (nodes.py:134-144).

    Above the fraction (or with the feature off) the count is unchanged; below
    it the count ramps down as round(n * frac / min_frac) with Python's
    half-to-even rounding, never below zero."""
    if min_frac <= 0 or frac >= min_frac or n_steps <= 0:
        return n_steps
    return max(0, round(n_steps * frac / min_frac))


def effective_inner_steps(n_steps: int, sigmas_host: Sequence[float], sigma_mean: float, abt_frac: float,
                          early_stop: int = 1, min_frac: float = 1.0) -> int:
    """n_eff of nodes.py:286-299 from host values.

    `abt_frac` must be float(float32 mean(1 - abt)) so the banker's rounding sees
    the same number the reference does."""
    diffs = [abs(s - sigma_mean) for s in sigmas_host]
    current = diffs.index(min(diffs))  # torch.argmin returns the first minimum, so does index()
    total = len(sigmas_host) - 1
    if total - current <= early_stop:
        return 0
    return min_step_frac_effective_steps(n_steps, abt_frac, min_frac)


@dataclass
class Hyper:
    step_size: float
    lam: float
    beta: float
    min_step_frac: float
    flow: bool

    def to_c(self) -> _native.Hyper:
        return _native.Hyper(float(self.step_size), float(self.lam), float(self.beta),
                             float(self.min_step_frac), 1 if self.flow else 0, 0)


def build_table(abt: Sequence[float], ve_sigma: Sequence[float], hyper: Hyper,
                rep_noise: Optional[Sequence[float]] = None, rep_y: Optional[Sequence[float]] = None,
                corr: Optional[Sequence[float]] = None, out: Optional[np.ndarray] = None) -> np.ndarray:
    """Coefficient table [rows, 24] fp32 (host) via the library's lp_build_coef_table."""
    lib = _native.load()
    a = np.ascontiguousarray(abt, dtype=np.float64)
    v = np.ascontiguousarray(ve_sigma, dtype=np.float64)
    n = a.shape[0]
    assert v.shape[0] == n
    if out is None:
        out = np.empty((n, TABLE_STRIDE), dtype=np.float32)
    assert out.dtype == np.float32 and out.size >= n * TABLE_STRIDE and out.flags["C_CONTIGUOUS"]

    def opt(x):
        if x is None:
            return None, None
        arr = np.ascontiguousarray(x, dtype=np.float64)
        assert arr.shape[0] == n
        return arr, arr.ctypes.data_as(C.c_void_p)

    rn, rn_p = opt(rep_noise)
    ry, ry_p = opt(rep_y)
    cc, cc_p = opt(corr)
    hc = hyper.to_c()
    rc = lib.lp_build_coef_table(a.ctypes.data_as(C.c_void_p), v.ctypes.data_as(C.c_void_p), rn_p, ry_p, cc_p,
                                 n, C.byref(hc), out.ctypes.data_as(C.c_void_p))
    _native.check(rc, "lp_build_coef_table")
    return out


def mean_half_dt(abt: Sequence[float], hyper: Hyper) -> float:
    """mean over the batch of dtx/2 = StepSize*clamp(1-abt, MinStepFrac); the reference
    skips the whole sub-step when it is <= 0 (lanpaint.py:205)."""
    vals = [hyper.step_size * max(1.0 - float(a), hyper.min_step_frac) for a in abt]
    return sum(vals) / max(1, len(vals))
