"""Host logic of the node layer's engine / graph cache (no GPU): what the key distinguishes, what it ignores, when
an entry is dropped.  A captured graph bakes in everything the key is made of."""
import types

import pytest
import torch

import minicomfy

minicomfy.install()
from lanpaint_b200 import comfy_nodes as N  # noqa: E402
from lanpaint_b200.engine import options_fingerprint  # noqa: E402


def _guider(patcher, pos=0.3, neg=-0.2, cfg=5.0, cfg_big=5.0):
    g = minicomfy.CFGGuider(patcher)
    g.set_conds(pos, neg)
    g.set_cfg(cfg)
    g.cfg_BIG = cfg_big
    g.inner_model = patcher.model
    return g


HYPER = dict(NSteps=5, Friction=15.0, Lambda=5.0, Beta=1.0, StepSize=0.2, IS_FLUX=False, IS_FLOW=False,
             EarlyStopThreshold=0.0, EarlyStopPatience=1, MinStepFrac=1.0)
SIG = [14.6, 7.0, 3.0, 1.0, 0.0]


def _lookup(cache, guider, x, opts=None, model_options=None, sig=SIG, hyper=HYPER, has_cb=True, max_denoise=True):
    return cache.lookup(guider, x, sig, hyper, 1, max_denoise, opts or {}, model_options or {}, has_cb)


def test_options_fingerprint_is_structural():
    a = {"transformer_options": {"sample_sigmas": torch.tensor([1.0, 0.5]), "k": 1}, "x": [1, 2]}
    b = {"x": [1, 2], "transformer_options": {"k": 1, "sample_sigmas": torch.tensor([1.0, 0.5])}}
    assert options_fingerprint(a) == options_fingerprint(b)                 # order and tensor identity do not matter
    b["transformer_options"]["sample_sigmas"] = torch.tensor([1.0, 0.25])
    assert options_fingerprint(a) != options_fingerprint(b)                 # small tensors count by content
    big = torch.zeros(4096)
    assert options_fingerprint({"p": big}) == options_fingerprint({"p": big})
    assert options_fingerprint({"p": big}) != options_fingerprint({"p": torch.zeros(4096)})   # large ones by identity
    f = lambda a: a  # noqa: E731
    assert options_fingerprint({"hook": [f]}) == options_fingerprint({"hook": [f]})


def test_cache_key_distinguishes_what_a_graph_bakes_in():
    cache = N._EngineCache(capacity=32)
    patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(lambda x, s, c: x), "cpu")
    x = torch.zeros(2, 4, 8, 8)
    pos, neg = object(), object()          # CONDITIONING objects: identity is what ComfyUI's cache preserves
    e = _lookup(cache, _guider(patcher, pos, neg), x)
    assert _lookup(cache, _guider(patcher, pos, neg), x) is e                # a new guider object per call: same entry
    assert _lookup(cache, _guider(patcher, pos, neg, cfg=7.0), x) is not e
    assert _lookup(cache, _guider(patcher, pos, neg, cfg_big=-0.5), x) is not e
    assert _lookup(cache, _guider(patcher, object(), neg), x) is not e
    assert _lookup(cache, _guider(patcher, pos, neg), torch.zeros(1, 4, 8, 8)) is not e
    assert _lookup(cache, _guider(patcher, pos, neg), x, sig=[14.6, 7.0, 0.0]) is not e
    assert _lookup(cache, _guider(patcher, pos, neg), x, hyper=dict(HYPER, Lambda=8.0)) is not e
    assert _lookup(cache, _guider(patcher, pos, neg), x, opts={"rng": "philox"}) is not e
    assert _lookup(cache, _guider(patcher, pos, neg), x, model_options={"transformer_options": {"patch": 1}}) is not e
    assert _lookup(cache, _guider(patcher, pos, neg), x, has_cb=False) is not e
    assert _lookup(cache, _guider(patcher, pos, neg), x, max_denoise=False) is not e
    other = minicomfy.ModelPatcher(minicomfy.BaseModel(lambda x, s, c: x), "cpu")
    assert _lookup(cache, _guider(other, pos, neg), x) is not e
    assert _lookup(cache, _guider(patcher, pos, neg), x) is e                # and the original is still there
    # the engine's own options live under "lanpaint_b200" and are keyed separately from what the network reads
    assert _lookup(cache, _guider(patcher, pos, neg), x, model_options={"lanpaint_b200": {"timing": True}}) is e


def test_conditioning_is_keyed_by_what_it_wraps_not_by_the_per_call_containers():
    """ComfyUI's CFGGuider.set_conds -> convert_cond builds a new list of new dicts (fresh uuid, new CONDCrossAttn
    wrappers) around the same text-encoder tensors on every call: that must hit the same entry; another prompt
    (other tensors), another strength or an extra cond entry must not."""
    import uuid

    class CONDCrossAttn:       # comfy.conds.*: a thin wrapper with one `.cond`
        def __init__(self, cond):
            self.cond = cond

    def convert(cross, pooled, strength=1.0):
        return [{"cross_attn": cross, "pooled_output": pooled, "strength": strength, "uuid": uuid.uuid4(),
                 "model_conds": {"c_crossattn": CONDCrossAttn(cross)}}]
    cache = N._EngineCache(capacity=32)
    patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(lambda x, s, c: x), "cpu")
    x = torch.zeros(1, 4, 8, 8)
    t_pos, t_neg = torch.zeros(1, 77, 16), torch.zeros(1, 77, 16)
    p_pos, p_neg = torch.zeros(1, 8), torch.zeros(1, 8)
    e = _lookup(cache, _guider(patcher, convert(t_pos, p_pos), convert(t_neg, p_neg)), x)
    assert _lookup(cache, _guider(patcher, convert(t_pos, p_pos), convert(t_neg, p_neg)), x) is e
    assert _lookup(cache, _guider(patcher, convert(torch.zeros(1, 77, 16), p_pos), convert(t_neg, p_neg)), x) is not e
    assert _lookup(cache, _guider(patcher, convert(t_pos, p_pos, 0.8), convert(t_neg, p_neg)), x) is not e
    assert _lookup(cache, _guider(patcher, convert(t_pos, p_pos) * 2, convert(t_neg, p_neg)), x) is not e
    assert _lookup(cache, _guider(patcher, convert(t_pos, p_pos), convert(t_neg, p_neg)), x) is e
    assert e.keep[2][0]["cross_attn"] is t_pos          # the entry pins what its ids name


def test_entry_is_dropped_when_the_weights_move_and_when_the_cache_is_full():
    cache = N._EngineCache(capacity=2)
    net = torch.nn.Linear(4, 4)
    patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(net), "cpu")
    x = torch.zeros(1, 4, 8, 8)
    pos, neg = object(), object()
    e = _lookup(cache, _guider(patcher, pos, neg), x)
    e.runs = 3
    assert _lookup(cache, _guider(patcher, pos, neg), x) is e
    with torch.no_grad():
        net.weight.mul_(2.0)                                   # same storage: graphs read weights by address
    assert _lookup(cache, _guider(patcher, pos, neg), x) is e
    net.weight.data = net.weight.data.clone()                  # new storage (a re-loaded model): graphs are dead
    fresh = _lookup(cache, _guider(patcher, pos, neg), x)
    assert fresh is not e and fresh.runs == 0
    a = _lookup(cache, _guider(patcher, pos, neg, cfg=1.5), x)
    b = _lookup(cache, _guider(patcher, pos, neg, cfg=2.5), x)   # capacity 2: the oldest entry goes
    assert len(cache.entries) == 2 and _lookup(cache, _guider(patcher, pos, neg, cfg=2.5), x) is b
    assert _lookup(cache, _guider(patcher, pos, neg), x) is not fresh
    assert a is not b


def test_graph_switches():
    import os
    assert N._graphs_enabled({}) and not N._graphs_enabled({"cuda_graph": False})
    os.environ["LANPAINT_B200_GRAPH"] = "0"
    try:
        assert not N._graphs_enabled({})
    finally:
        del os.environ["LANPAINT_B200_GRAPH"]
