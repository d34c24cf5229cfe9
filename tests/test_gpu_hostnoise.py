"""GPU: ComfyUI's CPU noise image drawn on the device (csrc/lp_hostnoise.cu, lanpaint_b200/hostnoise.py).

`lp_torch_cpu_randn_f32` must produce the bits of `torch.manual_seed(seed); torch.randn(size, device="cpu")` -- the
oracle here is torch itself on this host -- for whole 16-groups, redrawn tails, seeds above 2^32, and sizes that
span many twists of the generator; the wrapper must leave torch's CPU and CUDA generators where ComfyUI's call leaves
them; and a node call that draws its noise on the device must return the latent of one that lets ComfyUI draw it."""
import functools
import sys
import time

import pytest
import torch

import minicomfy

# torch's CPU randn draws the avx_mathfun bits on every x86 host with AVX2 (the AVX512 dispatch falls back to that
# kernel); on anything else the device draw is never enabled (hostnoise.verified) and there is nothing to compare
_AVX2_HOST = torch.backends.cpu.get_cpu_capability() in ("AVX2", "AVX512")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not _AVX2_HOST, reason="this host's torch.randn is not the AVX2 stream")]


def _cpu(shape, seed):
    return torch.randn(shape, generator=torch.Generator().manual_seed(seed), dtype=torch.float32, device="cpu")


def test_this_hosts_torch_draws_the_kernels_bits(cuda_device):
    from lanpaint_b200 import hostnoise
    assert hostnoise.verified(cuda_device), ("torch.randn on this host is not the AVX2 avx_mathfun stream: the node "
                                             "layer would keep ComfyUI's CPU draw")


@pytest.mark.parametrize("shape,seed", [((16,), 0), ((64,), 1), ((1000,), 2), ((4104,), 0xFFFFFFFF), ((1, 4, 128, 128), 7),
                                        ((3, 5, 7, 11), 2 ** 32 + 5), ((1 << 20,), 2 ** 63 + 11), ((624 * 227 + 16,), 3),
                                        ((8, 4, 128, 128), 1234567890123456789)])
def test_device_stream_is_torchs_cpu_randn_bit_for_bit(shape, seed, cuda_device):
    from lanpaint_b200 import hostnoise
    got = hostnoise.torch_cpu_randn(shape, seed, cuda_device, advance_cpu_generator=False)
    want = _cpu(shape, seed)
    assert got.shape == want.shape and got.is_cuda and got.is_contiguous()
    assert torch.equal(got.cpu(), want)


def test_batch_of_128_sdxl_latents_and_its_cost(cuda_device):
    from lanpaint_b200 import hostnoise
    shape = (128, 4, 128, 128)
    t0 = time.perf_counter()
    want = _cpu(shape, 99)
    t_cpu = time.perf_counter() - t0
    hostnoise.torch_cpu_randn(shape, 98, cuda_device)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    got = hostnoise.torch_cpu_randn(shape, 99, cuda_device)
    torch.cuda.synchronize()
    t_dev = time.perf_counter() - t0
    assert torch.equal(got.cpu(), want)
    print(f"\n8.4 M normals: torch.randn on the CPU {1e3 * t_cpu:.1f} ms, on the device {1e3 * t_dev:.2f} ms", file=sys.stderr)
    assert t_dev < t_cpu


def test_generators_are_left_where_comfyui_leaves_them(cuda_device):
    """comfy.sample.prepare_noise = torch.manual_seed(seed) + a draw from the default CPU generator: afterwards the
    CPU generator has consumed the draw and every CUDA generator is freshly seeded."""
    from lanpaint_b200 import hostnoise
    shape, seed = (2, 4, 33, 31), 4242                    # 8184 values: not a multiple of 16
    torch.manual_seed(seed)
    want = torch.randn(shape)
    cpu_state, cuda_state = torch.get_rng_state(), torch.cuda.get_rng_state(cuda_device)
    torch.manual_seed(1)
    torch.rand(3, device=cuda_device)
    got = hostnoise.torch_cpu_randn(shape, seed, cuda_device)
    assert torch.equal(got.cpu(), want)
    assert torch.equal(torch.get_rng_state(), cpu_state)
    assert torch.equal(torch.cuda.get_rng_state(cuda_device), cuda_state)
    assert torch.equal(torch.randn(100), torch.randn(100, generator=torch.Generator().set_state(cpu_state)))


def _denoiser(x, sigma, cond):
    return 0.7 * x + 0.1 * torch.tanh(x) + cond


@pytest.mark.parametrize("node", ["LanPaint_KSampler", "LanPaint_SamplerCustom"])
def test_node_call_is_unchanged_by_where_the_noise_is_drawn(node, cuda_device):
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    dev = cuda_device
    g = torch.Generator().manual_seed(3)
    y = torch.randn(2, 4, 32, 32, generator=g)
    noise_mask = (torch.rand(2, 1, 32, 32, generator=g) < 0.5).float()
    outs, calls = {}, {}
    real = sys.modules["comfy.sample"].prepare_noise
    for device_noise in (True, False):
        n_cpu = {"n": 0}

        @functools.wraps(real)       # still ComfyUI's stock function as far as the patch layer can tell
        def counting(latent_image, seed, noise_inds=None):
            n_cpu["n"] += 1
            return real(latent_image, seed, noise_inds)
        sys.modules["comfy.sample"].prepare_noise = counting
        try:
            patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(_denoiser), dev)
            patcher.model_options["lanpaint_b200"] = {"device_noise": device_noise, "cuda_graph": False}
            latent = {"samples": y, "noise_mask": noise_mask}
            if node == "LanPaint_KSampler":
                (out,) = N.LanPaint_KSampler().sample(patcher, 31, 8, 5.0, "euler", "karras", 0.3, -0.2, latent, 1.0, 3,
                                                      "Image First", "", N.IMAGE_MODE)
            else:
                out, _ = N.LanPaint_SamplerCustom().sample(patcher, minicomfy.ksampler("euler"),
                                                           minicomfy.get_sigmas_karras(8, 0.0292, 14.6146), True, 31, 5.0,
                                                           0.3, -0.2, latent, 3, "Image First", "")
        finally:
            sys.modules["comfy.sample"].prepare_noise = real
        outs[device_noise], calls[device_noise] = out["samples"], n_cpu["n"]
    assert torch.equal(outs[True], outs[False])
    if node == "LanPaint_KSampler":
        assert calls[True] == 0 and calls[False] == 1     # the CPU draw really was replaced, and really is the fallback
    assert sys.modules["comfy.sample"].prepare_noise is real


def test_what_the_kernel_does_not_cover_goes_to_comfyui(cuda_device):
    """batch_index noise (one draw per index), non-fp32 latents and tiny latents keep ComfyUI's own function."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    seen = []
    real = sys.modules["comfy.sample"].prepare_noise
    spy = functools.wraps(real)(lambda *a, **k: (seen.append(a[0].shape), real(*a, **k))[1])
    sys.modules["comfy.sample"].prepare_noise = spy
    try:
        with N.override_sample_function(cuda_device):
            fn = sys.modules["comfy.sample"].prepare_noise
            a = fn(torch.zeros(1, 4, 8, 8), 5, [0, 2])
            b = fn(torch.zeros(1, 4, 8, 8, dtype=torch.float64), 5)
            c = fn(torch.zeros(1, 1, 2, 2), 5)
            d = fn(torch.zeros(1, 4, 8, 8), 5)
        assert len(seen) == 3 and a.device.type == b.device.type == c.device.type == "cpu" and d.is_cuda
        assert torch.equal(d.cpu(), real(torch.zeros(1, 4, 8, 8), 5))
    finally:
        sys.modules["comfy.sample"].prepare_noise = real


def test_a_foreign_noise_source_is_left_alone(cuda_device):
    """An extension that replaced comfy.sample.prepare_noise (GPU noise, other RNGs ...) keeps working inside a
    LanPaint node: only ComfyUI's stock function is swapped for the device draw."""
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    real = sys.modules["comfy.sample"].prepare_noise

    def their_noise(latent_image, seed, noise_inds=None):
        return torch.full_like(latent_image, 0.25)
    sys.modules["comfy.sample"].prepare_noise = their_noise
    try:
        with N.override_sample_function(cuda_device):
            assert sys.modules["comfy.sample"].prepare_noise is their_noise
    finally:
        sys.modules["comfy.sample"].prepare_noise = real
