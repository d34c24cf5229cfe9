"""The written-down algorithm of the device-side noise image (oracle/torch_cpu_randn.py: the numpy restatement of
lp_hostnoise.cu) against torch itself, on the CPU: generator, uniform conversion, 16-group Box-Muller with the redrawn
tail, and the FMA-contraction pattern of torch's AVX2 log / sincos.  A host whose torch does not take the AVX2 path
draws other bits; there the device draw is never enabled (hostnoise.verified) and this pin is skipped."""
import numpy as np
import pytest
import torch

from oracle import torch_cpu_randn as R


def _host_draws_the_avx2_stream():
    want = torch.randn(64, generator=torch.Generator().manual_seed(1)).numpy()
    got, _ = R.torch_cpu_randn(1, 64)
    return np.array_equal(want, got)


def test_generator_and_uniform_conversion_are_torchs():
    for seed in (0, 7, 2 ** 32 + 17, 2 ** 63 + 5):
        want = torch.rand(3000, generator=torch.Generator().manual_seed(seed)).numpy()
        assert np.array_equal(R.uniforms(seed, 3000), want), seed          # seeds are truncated to 32 bits


@pytest.mark.parametrize("seed,n", [(0, 16), (1, 64), (12345, 1000), (2 ** 32 + 17, 4104), (2 ** 63 + 5, 65536),
                                    (99, 624 * 5 + 8)])
def test_normals_are_torchs_bit_for_bit(seed, n):
    if not _host_draws_the_avx2_stream():
        pytest.skip("torch.randn on this host does not take the AVX2 avx_mathfun path")
    gen = torch.Generator().manual_seed(seed)
    want = torch.randn(n, generator=gen).numpy()
    got, consumed = R.torch_cpu_randn(seed, n)
    assert np.array_equal(got, want)
    # the generator has moved by exactly `consumed` outputs: its next uniform is output number `consumed`
    nxt = torch.rand(4, generator=gen).numpy()
    assert np.array_equal(nxt, R.uniforms(seed, consumed + 4)[consumed:])


def test_shaped_draw_is_the_flat_stream():
    if not _host_draws_the_avx2_stream():
        pytest.skip("torch.randn on this host does not take the AVX2 avx_mathfun path")
    want = torch.randn((3, 4, 16, 16), generator=torch.Generator().manual_seed(5)).numpy().reshape(-1)
    assert np.array_equal(R.torch_cpu_randn(5, want.size)[0], want)
