"""lanpaint_b200: B200-native drop-in for LanPaint's inner Langevin loop.

Public surface
  LanPaint            the engine seam (reference: src/LanPaint/lanpaint.py)
  LangevinState       reference: src/LanPaint/types.py  (here: lanpaint_b200/state.py)
  NODE_CLASS_MAPPINGS / NODE_DISPLAY_NAME_MAPPINGS / WEB_DIRECTORY
                      the ComfyUI custom-node protocol (reference: __init__.py:90-98,
                      src/LanPaint/nodes.py:1347-1378), resolved lazily because they need ComfyUI; where
                      ComfyUI is absent (node-diff CI, unit tests) a tooling stub of its module surface
                      is installed first so the classes can be introspected, like the reference's
                      __init__.py:14-98 does
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
        try:
            from . import comfy_nodes
        except ModuleNotFoundError as e:
            if e.name not in ("comfy", "nodes", "latent_preview", "comfyui_version"):
                raise
            # no ComfyUI here (node-diff CI, unit tests): the classes can still be imported and introspected
            # (reference __init__.py:90-96 does the same); sampling needs the real thing
            from . import _tooling_stubs
            _tooling_stubs.install()
            from . import comfy_nodes
        return getattr(comfy_nodes, name)
    raise AttributeError(name)
