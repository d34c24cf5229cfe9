"""lanpaint_b200: B200-native drop-in for LanPaint's inner Langevin loop.

Public surface
  LanPaint            the engine seam (reference: src/LanPaint/lanpaint.py)
  LangevinState       reference: src/LanPaint/types.py  (here: lanpaint_b200/state.py)
  NODE_CLASS_MAPPINGS / NODE_DISPLAY_NAME_MAPPINGS / WEB_DIRECTORY
                      the ComfyUI custom-node protocol (reference: __init__.py:90-98,
                      src/LanPaint/nodes.py:1347-1378), resolved lazily because they need ComfyUI
"""
from .state import LangevinState  # noqa: F401

__version__ = "0.1.0"
WEB_DIRECTORY = "./web"

_LAZY = {"NODE_CLASS_MAPPINGS", "NODE_DISPLAY_NAME_MAPPINGS"}


def __getattr__(name):
    if name == "LanPaint":
        from .engine import LanPaint
        return LanPaint
    if name == "NoiseTape":
        from .engine import NoiseTape
        return NoiseTape
    if name in _LAZY:
        from . import comfy_nodes
        return getattr(comfy_nodes, name)
    raise AttributeError(name)
