"""Shared by tests/golden/make_node_golden.py (which drives the REFERENCE's node layer) and the node-level
parity tests (which drive lanpaint_b200's): how a case's JSON arguments become the positional call ComfyUI's
executor makes.  Test infrastructure."""
import torch

import minicomfy

IMAGE_MODE = "🖼️ Image Inpainting"
VIDEO_MODE = "🎬 Video Inpainting"


def denoiser(x, sigma, cond):
    """pointwise cond-dependent x0 prediction; `cond` is a float standing in for CONDITIONING"""
    return 0.7 * x + 0.1 * torch.tanh(x) + cond


class FixedNoise:
    """a NOISE object for LanPaint_SamplerCustomAdvanced (ComfyUI's RandomNoise protocol: .seed, .generate_noise)"""

    def __init__(self, seed, image):
        self.seed, self.image = seed, image

    def generate_noise(self, latent):
        return self.image


POS, NEG = 0.3, -0.2


def build_patcher(c, device="cpu", net=denoiser):
    mtype = getattr(minicomfy.ModelType, c.get("model_type", "EPS"))
    cls = minicomfy.WAN22 if c.get("wan22") else minicomfy.BaseModel
    base = cls(net, model_type=mtype, latent_channels=c["shape"][1], shift=c.get("shift", 1.0))
    return minicomfy.ModelPatcher(base, device)


def call_node(N, c, patcher, latent, noise_image=None):
    """The same positional call ComfyUI's executor makes, on node module `N` (the reference's or lanpaint_b200's)."""
    a = c["args"]
    node = N.NODE_CLASS_MAPPINGS[c["node"]]()
    if c["node"] == "LanPaint_KSampler":
        return node.sample(patcher, a["seed"], a["steps"], a["cfg"], a["sampler_name"], a["scheduler"], POS, NEG, latent,
                           a["denoise"], a["LanPaint_NumSteps"], a["LanPaint_PromptMode"], "", a["Inpainting_mode"])
    if c["node"] == "LanPaint_KSamplerAdvanced":
        return node.sample(patcher, a["add_noise"], a["noise_seed"], a["steps"], a["cfg"], a["sampler_name"],
                           a["scheduler"], POS, NEG, latent, a["start_at_step"], a["end_at_step"],
                           a["return_with_leftover_noise"], a["LanPaint_NumSteps"], a["LanPaint_Lambda"],
                           a["LanPaint_StepSize"], a["LanPaint_PromptMode"], "", a["Inpainting_mode"])
    kind, n = a["sigmas"]
    sig = minicomfy.get_sigmas_karras(n, 0.0292, 14.6146)
    sampler = minicomfy.ksampler(a["sampler"])
    if c["node"] == "LanPaint_SamplerCustom":
        return node.sample(patcher, sampler, sig, a["add_noise"], a["noise_seed"], a["cfg"], POS, NEG, latent,
                           a["LanPaint_NumSteps"], a["LanPaint_PromptMode"], "")
    guider = minicomfy.CFGGuider(patcher)
    guider.set_conds(POS, NEG)
    guider.set_cfg(a["cfg"])
    return node.sample(FixedNoise(a["noise_seed"], noise_image), guider, sampler, sig, latent, a["LanPaint_NumSteps"],
                       a["LanPaint_Lambda"], a["LanPaint_StepSize"], a["LanPaint_PromptMode"], "")
