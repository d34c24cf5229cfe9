"""Inner-loop early stop for the B200 engine (reference: src/LanPaint/earlystop.py).

Same decisions as `LanPaintEarlyStopper` -- threshold scaled by 4*abt*(1-abt), patience+1 consecutive
stable checks, drift guard against an anchor, optional user `distance_fn`, optional trace list -- but the
metric itself (two weighted MSEs of consecutive x0 estimates over the inpaint region and over its
4-neighbour boundary ring, earlystop.py:32-55) is ONE reduction kernel (`lp_stop_stats_f32`: warp shuffles,
one atomic per block) and ONE read-back per check where the reference issues 2-4 reductions and `.item()`s.

Off unless `EarlyStopThreshold > 0` or `model_options["lanpaint_semantic_stop"]` asks for it; every sampler
node passes threshold 0.0 (src/LanPaint/nodes.py:499,575,693,776).
"""
from __future__ import annotations

import ctypes as C
import inspect
from typing import Any, Callable, Optional

import torch

from . import _native

_P = C.c_void_p


def _clamp01(v: float) -> float:
    return 0.0 if v <= 0.0 else (1.0 if v >= 1.0 else v)


def abt_scale(abt_val: float) -> float:
    """0 at abt = 0 or 1, 1 at abt = 0.5 (earlystop.py:21-29)."""
    a = _clamp01(abt_val)
    return _clamp01(4.0 * a * (1.0 - a))


def boundary_ring(mask_u8: torch.Tensor) -> Optional[torch.Tensor]:
    """uint8 ring = unknown sites with a known 4-neighbour (earlystop.py:32-48); 4-D latents only."""
    if mask_u8.dim() != 4:
        return None
    known = mask_u8 != 0
    near = torch.zeros_like(known)
    near[:, :, 1:, :] |= known[:, :, :-1, :]
    near[:, :, :-1, :] |= known[:, :, 1:, :]
    near[:, :, :, 1:] |= known[:, :, :, :-1]
    near[:, :, :, :-1] |= known[:, :, :, 1:]
    return ((~known) & near).to(torch.uint8).contiguous()


def _wrap_distance_fn(fn: Optional[Callable[..., Any]]):
    """Normalise a user hook to fn(prev, cur, ctx) -> scalar | None (earlystop.py:184-236):
    3+ positional / *args -> (prev, cur, ctx); a `ctx` keyword or **kw -> (prev, cur, ctx=ctx);
    otherwise the legacy two-argument order (cur, prev)."""
    if not callable(fn):
        return None
    try:
        params = list(inspect.signature(fn).parameters.values())
    except (ValueError, TypeError):
        def fallback(p, c, ctx):
            try:
                return fn(p, c, ctx)
            except TypeError:
                return fn(c, p)
        return fallback
    kinds = inspect.Parameter
    positional = [p for p in params if p.kind in (kinds.POSITIONAL_ONLY, kinds.POSITIONAL_OR_KEYWORD)]
    if len(positional) >= 3 or any(p.kind == kinds.VAR_POSITIONAL for p in params):
        return lambda p, c, ctx: fn(p, c, ctx)
    if any(p.name == "ctx" for p in params) or any(p.kind == kinds.VAR_KEYWORD for p in params):
        return lambda p, c, ctx: fn(p, c, ctx=ctx)
    return lambda p, c, ctx: fn(c, p)


class EarlyStopper:
    def __init__(self, *, threshold, threshold_eff, patience_eff, mask, ring, w_inpaint, w_ring, dims, distance_fn,
                 trace, bench_ids, abt_val, device, reduce=None):
        self.enabled = True
        self.threshold = float(threshold)
        self.threshold_eff = float(threshold_eff)
        self.patience_eff = int(patience_eff)
        self.mask, self.ring = mask, ring
        self.w_inpaint, self.w_ring = float(w_inpaint), (None if w_ring is None else float(w_ring))
        self.dims = dims
        self.trace = trace
        self.bench_case_id, self.bench_outer_step, self.bench_timestep = bench_ids
        self.abt_val = abt_val
        self.patience_counter = 0
        self.x0_anchor: Optional[torch.Tensor] = None
        self._dist_wrapper = _wrap_distance_fn(distance_fn)
        self._sums = torch.zeros(2, dtype=torch.float64, device=device)
        # a latent sharded across ranks (frame_shard.py): the two masked sums are the ONLY cross-shard quantity of the
        # whole hot path; `reduce(tensor)` sums them over the shards in place (one all_reduce of 2 doubles per check)
        self._reduce = reduce
        self._bufs = [None, None]
        self._turn = 0
        self.reads = 0  # host read-backs issued (one per statistics kernel)

    @property
    def has_custom_distance_fn(self) -> bool:
        return self._dist_wrapper is not None

    def next_x0e_buffer(self, like: torch.Tensor) -> torch.Tensor:
        k = self._turn & 1
        self._turn += 1
        if self._bufs[k] is None or self._bufs[k].shape != like.shape:
            self._bufs[k] = torch.empty_like(like)
        return self._bufs[k]

    def _local_sums(self, a: torch.Tensor, b: torch.Tensor, table=None) -> torch.Tensor:
        """This shard's two masked sums of squared differences, on the device (lp_stop_stats_f32)."""
        lib = _native.load()
        rc = lib.lp_stop_stats_f32(_P(a.data_ptr()), _P(b.data_ptr()), _P(self.mask.data_ptr()),
                                   _P(self.ring.data_ptr()) if self.ring is not None else None,
                                   _P(table.data_ptr()) if table is not None else None, C.byref(self.dims),
                                   _P(self._sums.data_ptr()), _P(torch.cuda.current_stream(a.device).cuda_stream))
        _native.check(rc, "lp_stop_stats_f32")
        return self._sums

    def _stats(self, a: torch.Tensor, b: torch.Tensor, table=None):
        """(weighted MSE over the inpaint region, over the ring or None) -- earlystop.py:51-55."""
        sums = self._local_sums(a, b, table)
        if self._reduce is not None:
            self._reduce(sums)
        s_in, s_ring = sums.tolist()  # the one host read-back of this check
        self.reads += 1
        d_in = s_in / (self.w_inpaint + 1e-12)
        d_ring = None if self.ring is None else s_ring / (self.w_ring + 1e-12)
        return d_in, d_ring

    def step(self, *, i: int, n_steps: int, x_before: Optional[torch.Tensor], x_after: torch.Tensor,
             x0_prev: Optional[torch.Tensor], x0_cur: Optional[torch.Tensor], table: torch.Tensor,
             custom_prev=None, custom_cur=None, ctx: Optional[dict] = None) -> bool:
        """One check after sub-step i (earlystop.py:238-336).  x_before/x_after are model-space states
        (the table's inv_S turns their difference into the VP-space one); x0_* are the x_t+score buffers."""
        dist = None
        custom = False
        d_in = d_ring = d_drift = None
        if self._dist_wrapper is not None:
            dist = self._dist_wrapper(custom_prev() if callable(custom_prev) else custom_prev,
                                      custom_cur() if callable(custom_cur) else custom_cur, ctx or {})
            if dist is not None:
                if isinstance(dist, torch.Tensor):
                    if dist.numel() != 1:
                        raise TypeError("distance_fn must return None or a scalar / 0-d (1-element) tensor")
                    dist = float(dist.item())
                else:
                    dist = float(dist)
        custom = dist is not None
        have_x0 = x0_prev is not None and x0_cur is not None
        if dist is None:
            if have_x0:
                d_in, d_ring = self._stats(x0_cur, x0_prev)
                dist = d_in if d_ring is None else max(d_in, d_ring)
            else:
                d_in, _ = self._stats(x_after, x_before, table)
                dist = d_in
        used = self.threshold if custom else self.threshold_eff

        if x0_cur is not None and not custom:  # drift guard, default metric only
            if dist <= used:
                if self.x0_anchor is None:
                    self.x0_anchor = x0_cur.clone()
                else:
                    a_in, a_ring = self._stats(x0_cur, self.x0_anchor)
                    d_drift = a_in if a_ring is None else max(a_in, a_ring)
                    dist = max(dist, d_drift)
            else:
                self.x0_anchor = None

        if dist <= used:
            self.patience_counter += 1
        else:
            self.patience_counter = 0
            self.x0_anchor = None
        stop = self.patience_counter >= self.patience_eff

        if isinstance(self.trace, list):
            self.trace.append({
                "case_id": self.bench_case_id, "outer_step": self.bench_outer_step,
                "bench_timestep": self.bench_timestep, "inner_step": i + 1, "dist": dist,
                "dist_inpaint": None if d_in is None else float(d_in),
                "dist_ring": None if d_ring is None else float(d_ring),
                "dist_drift": None if d_drift is None else float(d_drift),
                "threshold": float(used), "threshold_eff": float(self.threshold_eff),
                "patience_counter": int(self.patience_counter), "patience_eff": int(self.patience_eff),
                "abt": None if self.abt_val is None else float(self.abt_val),
                "custom_dist": bool(custom), "stopped": bool(stop),
            })
        return bool(stop)


def make_stopper(*, model_options, default_threshold, default_patience, default_distance_fn, packed_mask, like,
                 abt_mean: float, dims, reduce=None) -> Optional[EarlyStopper]:
    """LanPaintEarlyStopper.from_options (earlystop.py:63-151) over a packed uint8 mask."""
    semantic = model_options.get("lanpaint_semantic_stop") if isinstance(model_options, dict) else None
    threshold = float(default_threshold)
    patience = int(default_patience)
    distance_fn = default_distance_fn
    if isinstance(semantic, dict):
        threshold = float(semantic.get("threshold", threshold))
        patience = int(semantic.get("patience", patience))
        distance_fn = semantic.get("distance_fn", distance_fn)
        if patience > 0 and semantic.get("min_steps") is not None:  # legacy knob -> patience floor
            try:
                ms = int(semantic.get("min_steps"))
            except (TypeError, ValueError):
                ms = 0
            if ms > 1:
                patience = max(patience, ms - 1)
    if not (threshold > 0.0 and patience > 0):
        return None
    threshold_eff = threshold * abt_scale(float(abt_mean))
    if threshold_eff <= 0.0:
        return None
    m = packed_mask.data
    channels = like.shape[1] if packed_mask.channel_stride == 0 else 1
    ring = boundary_ring(m) if like.dim() == 4 else None
    counts = torch.stack([(m == 0).sum(), ring.sum() if ring is not None else m.new_zeros(()).sum()]).to(torch.float64)
    if reduce is not None:   # weights of the whole latent, not of this shard
        reduce(counts)
    n_in, n_ring = counts.tolist()
    w_in = float(n_in) * channels
    if w_in < 1e-6:
        return None
    w_ring = None if ring is None else float(n_ring) * channels
    trace = model_options.get("lanpaint_semantic_trace") if isinstance(model_options, dict) else None
    ids = (None, None, None)
    if isinstance(trace, list) and isinstance(model_options, dict):
        ids = (model_options.get("bench_case_id"), model_options.get("bench_outer_step"),
               model_options.get("bench_timestep"))
    return EarlyStopper(threshold=threshold, threshold_eff=threshold_eff, patience_eff=max(1, patience) + 1, mask=m,
                        ring=ring, w_inpaint=w_in, w_ring=w_ring, dims=dims, distance_fn=distance_fn, trace=trace,
                        bench_ids=ids, abt_val=float(abt_mean), device=like.device, reduce=reduce)
