"""The state a Langevin sub-step hands to the next one.

Reference: `LangevinState(v, C, x0)` in src/LanPaint/types.py:6-9, plus the legacy tuple forms
`(v, C)` / `(v, C, x0)` that `langevin_dynamics` still accepts (src/LanPaint/lanpaint.py:193-199).
On the device the whole state of the fused loop is the `C` buffer; this tuple only exists at the
`LanPaint.langevin_dynamics` entry point and for the early stopper.
"""
from __future__ import annotations

from typing import Any, NamedTuple, Optional

import torch


class LangevinState(NamedTuple):
    v: Optional[torch.Tensor]   # velocity of the disabled second-order scheme: always None
    C: Optional[torch.Tensor]   # drift constant of the last sub-step
    x0: Optional[torch.Tensor]  # x_t + score: what the early stopper compares between sub-steps

    @classmethod
    def coerce(cls, args: Any) -> Optional["LangevinState"]:
        """None, a LangevinState, or one of the legacy tuples -> LangevinState (or None)."""
        if args is None or isinstance(args, cls):
            return args
        if isinstance(args, tuple) and len(args) >= 2:
            return cls(args[0], args[1], args[2] if len(args) >= 3 else None)
        raise TypeError(f"cannot interpret {type(args).__name__} as a LangevinState")
