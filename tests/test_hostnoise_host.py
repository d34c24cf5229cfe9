"""Host logic of the device-side noise image (no GPU): where the default CPU generator must be left after
`hostnoise.torch_cpu_randn` -- at::mt19937 twists whole blocks of 624, so (state array, next, left) follow from the
number of outputs consumed -- checked against torch's own generator with numpy's MT19937 standing in for the kernel."""
import numpy as np
import pytest
import torch

from lanpaint_b200 import hostnoise as H


def _init_genrand(seed):
    mt = np.zeros(624, dtype=np.uint64)
    mt[0] = seed & 0xFFFFFFFF
    for i in range(1, 624):
        mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> np.uint64(30))) + i) & 0xFFFFFFFF
    return mt.astype(np.uint32)


def _engine_block_after(seed, consumed):
    bg = np.random.MT19937()
    bg.state = {"bit_generator": "MT19937", "state": {"key": _init_genrand(seed), "pos": 624}}
    bg.random_raw(((consumed + 623) // 624) * 624)
    return torch.from_numpy(bg.state["state"]["key"].astype(np.uint32).view(np.int32).copy())


@pytest.mark.parametrize("seed,n", [(77, 1000), (5, 64), (2 ** 40 + 3, 1248), (9, 623), (0, 16), (2 ** 64 - 1, 624 * 3 + 16)])
def test_cpu_generator_is_left_where_the_cpu_draw_leaves_it(seed, n):
    consumed = n + (16 if n % 16 else 0)          # a size that is not a multiple of 16 redraws its last 16 values
    g = torch.Generator()
    g.manual_seed(seed)
    H._advance_cpu_generator(g, _engine_block_after(seed, consumed), consumed)
    ref = torch.Generator()
    ref.manual_seed(seed)
    torch.randn(n, generator=ref)
    assert torch.equal(g.get_state(), ref.get_state())
    assert torch.equal(torch.randn(48, generator=g), torch.randn(48, generator=ref))
    assert torch.equal(torch.rand(5, generator=g, dtype=torch.float64), torch.rand(5, generator=ref, dtype=torch.float64))


def test_there_is_no_cpu_path():
    with pytest.raises(RuntimeError):
        H.torch_cpu_randn((4, 4), 1, "cpu")
    assert H.verified("cpu") is False


def test_what_the_node_layer_hands_to_the_kernel():
    """one fp32, strided draw of at least 16 values without batch_index noise; and never without a CUDA device"""
    import minicomfy
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    ok = torch.zeros(1, 4, 8, 8)
    assert N._device_randn_ok(ok) and not N._device_randn_ok(ok, [0, 1])
    assert not N._device_randn_ok(ok.double()) and not N._device_randn_ok(torch.zeros(1, 1, 3, 3))
    assert not N._device_randn_ok(None)
    patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(lambda x, s, c: x), "cpu")
    assert N._noise_device(patcher) is None                    # the model does not live on a GPU
    patcher.load_device = torch.device("cuda", 0)
    patcher.model_options["lanpaint_b200"] = {"device_noise": False}
    assert N._noise_device(patcher) is None                    # switched off: ComfyUI's own prepare_noise
    # with no verified device nothing is swapped beyond the reference's four functions
    import sys
    stock = sys.modules["comfy.sample"].prepare_noise
    with N.override_sample_function(None):
        assert sys.modules["comfy.sample"].prepare_noise is stock


def test_fifth_swap_is_scoped_like_the_other_four():
    """With a noise device the stock prepare_noise is swapped for the duration of the call only -- restored on normal
    exit, on an exception, and never captured as an "original" by a nested entry (nodes.py:384-421 semantics)."""
    import sys

    import minicomfy
    minicomfy.install()
    from lanpaint_b200 import comfy_nodes as N
    mod = sys.modules["comfy.sample"]
    stock = mod.prepare_noise
    dev = torch.device("cuda", 0)                 # only installed here, never called: no GPU needed
    with N.override_sample_function(dev):
        swapped = mod.prepare_noise
        assert swapped is not stock
        with N.override_sample_function(dev):     # nested entry: a no-op
            assert mod.prepare_noise is swapped
        assert mod.prepare_noise is swapped
        # what the kernel does not cover is handed to ComfyUI's own function even while swapped
        out = swapped(torch.zeros(1, 4, 4, 4), 3, [0])
        assert out.device.type == "cpu" and torch.equal(out, stock(torch.zeros(1, 4, 4, 4), 3, [0]))
    assert mod.prepare_noise is stock
    with pytest.raises(KeyError):
        with N.override_sample_function(dev):
            raise KeyError("boom")
    assert mod.prepare_noise is stock
