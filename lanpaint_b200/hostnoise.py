"""ComfyUI's CPU noise image, drawn on the GPU with the same bits.

`comfy.sample.prepare_noise(latent_image, seed)` -- called by `nodes.common_ksampler` for every KSampler-type node
(src/LanPaint/nodes.py:513,589) and by ComfyUI's RandomNoise objects -- is

    generator = torch.manual_seed(seed)
    torch.randn(latent_image.size(), dtype=..., generator=generator, device="cpu")

a single-threaded mt19937 + Box-Muller draw: 29 ms for a batch of 128 SDXL latents, three quarters of a LanPaint
KSampler call on a B200.  `lp_torch_cpu_randn_f32` (csrc/lp_hostnoise.cu) produces that stream on the device, bit for
bit, in about a millisecond; `torch_cpu_randn` below wraps it and leaves BOTH generators where ComfyUI's call would
have left them: `torch.manual_seed(seed)` is still issued (it also seeds every CUDA generator -- the Langevin draws of
the sampler run depend on that) and the default CPU generator is advanced by exactly the outputs the CPU draw would
have consumed.

The bits of torch's CPU `randn` are those of its AVX2 kernel (avx_mathfun's cephes log / sincos with FMA); a build
that dispatches elsewhere (no AVX2, another architecture) draws other bits.  So nothing is assumed: `verified(device)`
compares the device stream with this host's own `torch.randn` once per process and device, and the node layer only
replaces `prepare_noise` when they are identical.
"""
from __future__ import annotations

import ctypes as C
import struct
from typing import Dict, Sequence

import numpy as np
import torch

from . import _native

_MT_N = 624
_STATE_OFFSET = 24            # CPUGeneratorImplStateLegacy: u64 seed, i32 left, i32 seeded, u64 next, u64 state[624], ...
_verified: Dict[int, bool] = {}


def _numel(shape: Sequence[int]) -> int:
    n = 1
    for d in shape:
        n *= int(d)
    return n


def _draw(n: int, seed: int, device: torch.device):
    """-> (buffer of n + 16 floats whose first n are the normals, the engine's state words [624] on the device,
    number of generator outputs consumed)."""
    lib = _native.load()
    out = torch.empty(n + 16, dtype=torch.float32, device=device)
    state = torch.empty(_MT_N, dtype=torch.int32, device=device)
    consumed = C.c_int64(0)
    with torch.cuda.device(device):
        rc = lib.lp_torch_cpu_randn_f32(C.c_void_p(out.data_ptr()), n, C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF),
                                        C.c_void_p(state.data_ptr()), C.byref(consumed),
                                        C.c_void_p(torch.cuda.current_stream(device).cuda_stream))
    _native.check(rc, "lp_torch_cpu_randn_f32")
    return out, state, int(consumed.value)


def _advance_cpu_generator(gen: torch.Generator, state_words: torch.Tensor, consumed: int) -> None:
    """Put `gen` (freshly seeded) where `consumed` 32-bit draws would have left it: at::mt19937 twists a whole
    block of 624 at a time, so its state array is the last block generated (which the kernel completed) and
    next / left index into it."""
    blocks = (consumed + _MT_N - 1) // _MT_N
    nxt = consumed - _MT_N * (blocks - 1)
    left = _MT_N - nxt + 1
    raw = bytearray(gen.get_state().numpy().tobytes())
    struct.pack_into("<i", raw, 8, left)
    struct.pack_into("<Q", raw, 16, nxt)
    words = state_words.cpu().numpy().view(np.uint32).astype("<u8")
    raw[_STATE_OFFSET:_STATE_OFFSET + 8 * _MT_N] = words.tobytes()
    gen.set_state(torch.frombuffer(raw, dtype=torch.uint8).clone())


def torch_cpu_randn(shape: Sequence[int], seed: int, device, advance_cpu_generator: bool = True) -> torch.Tensor:
    """`torch.manual_seed(seed); torch.randn(shape, generator=<default CPU generator>, device="cpu")`, as a tensor on
    `device` with the same bits and the same effect on torch's generators.  fp32, at least 16 elements."""
    device = torch.device(device)
    n = _numel(shape)
    if device.type != "cuda":
        raise RuntimeError("lanpaint_b200.hostnoise draws on a CUDA device: there is no CPU path (torch.randn is one)")
    if n < 16:
        raise ValueError("torch's CPU randn takes another code path below 16 elements")
    gen = torch.manual_seed(seed)               # ComfyUI's own first line: CPU generator and every CUDA generator
    out, state, consumed = _draw(n, seed, device)
    if advance_cpu_generator:
        _advance_cpu_generator(gen, state, consumed)
    return out[:n].view(tuple(int(d) for d in shape))


def verified(device) -> bool:
    """True when the device stream equals THIS host's `torch.randn` (a body of whole 16-groups and a redrawn
    tail, two seeds, one of them above 2^32) and the generator bookkeeping reproduces the state torch's own draw
    leaves behind.  Checked once per process and device; never raises."""
    device = torch.device(device)
    key = device.index if device.index is not None else (torch.cuda.current_device() if device.type == "cuda" else -1)
    hit = _verified.get(key)
    if hit is not None:
        return hit
    ok = False
    try:
        if device.type == "cuda":
            ok = True
            for seed, n in ((0x5EED, 8192), (0x9E3779B97F4A7C15, 4096 + 8)):
                ref = torch.Generator().manual_seed(seed)
                want = torch.randn(n, generator=ref, dtype=torch.float32, device="cpu")
                got, state, consumed = _draw(n, seed, device)
                ok = ok and consumed == n + (16 if n % 16 else 0) and torch.equal(got[:n].cpu(), want)
                # ... and this torch's CPU generator state has the layout _advance_cpu_generator writes
                mine = torch.Generator().manual_seed(seed)
                _advance_cpu_generator(mine, state, consumed)
                ok = ok and torch.equal(mine.get_state(), ref.get_state())
    except Exception:
        ok = False
    _verified[key] = ok
    return ok
