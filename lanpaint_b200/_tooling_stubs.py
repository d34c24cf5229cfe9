"""Just enough of ComfyUI's module surface for the node CLASSES to be imported and introspected
(INPUT_TYPES / RETURN_TYPES / NODE_CLASS_MAPPINGS) where ComfyUI is not installed: ComfyUI's node-diff CI and
unit tests do exactly that with the reference (its __init__.py:14-98 installs an equivalent set).  Nothing here can
sample: every function a node calls while sampling lives in real ComfyUI modules that are deliberately absent."""
import sys
import types

_SCHEDULERS = ["normal", "karras", "exponential", "sgm_uniform", "simple", "ddim_uniform", "beta", "linear_quadratic",
               "kl_optimal", "AYS"]


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__lanpaint_b200_tooling_stub__ = True
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def install() -> bool:
    """Registers the stand-ins unless a `comfy` module is already importable.  Returns True if it did."""
    if "comfy" in sys.modules:
        return False
    try:
        import comfy  # noqa: F401  a real ComfyUI (or a test stand-in) on the path wins
        return False
    except ModuleNotFoundError:
        pass

    def repeat_to_batch_size(tensor, batch_size, dim=0):
        n = tensor.shape[dim]
        if n == batch_size:
            return tensor
        if n > batch_size:
            return tensor.narrow(dim, 0, batch_size)
        reps = [1] * tensor.ndim
        reps[dim] = -(-batch_size // n)
        return tensor.repeat(reps).narrow(dim, 0, batch_size)

    comfy = _module("comfy")
    comfy.__path__ = []
    comfy.utils = _module("comfy.utils", repeat_to_batch_size=repeat_to_batch_size, PROGRESS_BAR_ENABLED=False)
    comfy.samplers = _module("comfy.samplers", KSAMPLER=type("KSAMPLER", (), {}),
                             KSampler=type("KSampler", (), {"SCHEDULERS": list(_SCHEDULERS)}))
    comfy.model_base = _module("comfy.model_base", ModelType=types.SimpleNamespace(FLUX="FLUX", FLOW="FLOW"),
                               WAN22=type("WAN22", (), {}))
    sys.modules.setdefault("nodes", types.ModuleType("nodes"))
    sys.modules.setdefault("latent_preview", types.ModuleType("latent_preview"))
    if "comfyui_version" not in sys.modules:
        _module("comfyui_version", __version__="0.0.0")
    return True
