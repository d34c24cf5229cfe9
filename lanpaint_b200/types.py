"""State carried between Langevin sub-steps (reference: src/LanPaint/types.py:6-9)."""
from typing import NamedTuple, Optional

import torch


class LangevinState(NamedTuple):
    """(v, C, x0): v is always None in the live first-order scheme; C is the drift
    constant of the last sub-step; x0 is x_t + score, what the early stopper watches."""

    v: Optional[torch.Tensor]
    C: Optional[torch.Tensor]
    x0: Optional[torch.Tensor]
