"""One video latent sharded along its frame axis across ranks (BASELINE configs[4]).

Every quantity of the Langevin hot path is per element given per-sample scalars, so a sample [1,C,T,H,W] can be
cut along T (or any axis) and each rank can run the very same kernels on its slice with NO exchange.  The only
cross-shard quantities are scalars:
  * the early stopper's two masked sums of squared differences (reference: `_weighted_mse`,
    src/LanPaint/earlystop.py:51-55, used by `LanPaintEarlyStopper.step`, :238-313) and, once, the mask weights
    they are divided by -- one all_reduce of 2 doubles per check;
  * `mean(abs(noise)) < 1e-8` (lanpaint.py:51) -- known to the host.
This is the synthetic form of the configuration: the denoiser must be pointwise (or itself sequence-parallel).
A real video DiT attends across frames; sharding it needs sequence-parallel attention inside a third-party model,
which is outside this path -- with such a model every rank holds the whole sample ("replicas only").

`ShardGroup` is the small interface the run needs (rank, world, all_reduce_sum_, max_over_ranks, barrier);
`replicas.ReplicaGroup` provides it over torch.distributed (NCCL on the box, gloo in the CPU test) and
`ThreadGroup` provides it between threads of one process (one-GPU test of the whole sharded path).
"""
from __future__ import annotations

import threading
from typing import List, Optional, Sequence

import torch

from .replicas import shard_bounds


def frame_slice(n_frames: int, world: int, rank: int) -> slice:
    """Contiguous, balanced split of the frame axis (first n_frames % world ranks take one more)."""
    a, b = shard_bounds(n_frames, world, rank)
    return slice(a, b)


class DistGroup:
    """torch.distributed-backed group (wraps replicas.ReplicaGroup)."""

    def __init__(self, replica_group):
        self.g = replica_group
        self.rank, self.world = replica_group.rank, replica_group.world

    def all_reduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def max_over_ranks(self, v: float) -> float:
        return self.g.max_over_ranks(v)

    def barrier(self):
        self.g.barrier()


class ThreadGroup:
    """The same collective between `world` threads of one process: every participant adds its tensor into a shared
    accumulator, waits at a barrier, reads the total back.  Deterministic summation order (by rank)."""

    class _Shared:
        def __init__(self, world):
            self.world = world
            self.barrier = threading.Barrier(world)
            self.slots: List[Optional[torch.Tensor]] = [None] * world
            self.vals = [0.0] * world

    def __init__(self, shared: "ThreadGroup._Shared", rank: int):
        self.shared, self.rank, self.world = shared, rank, shared.world

    @classmethod
    def make(cls, world: int) -> List["ThreadGroup"]:
        sh = cls._Shared(world)
        return [cls(sh, r) for r in range(world)]

    def all_reduce_sum_(self, t: torch.Tensor) -> torch.Tensor:
        sh = self.shared
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()
        sh.slots[self.rank] = t.detach().clone()
        sh.barrier.wait()
        total = sh.slots[0].to(t.device).clone()
        for r in range(1, self.world):
            total += sh.slots[r].to(t.device)
        sh.barrier.wait()       # everyone has read the slots before anyone overwrites them
        t.copy_(total)
        return t

    def max_over_ranks(self, v: float) -> float:
        sh = self.shared
        sh.vals[self.rank] = float(v)
        sh.barrier.wait()
        out = max(sh.vals)
        sh.barrier.wait()
        return out

    def barrier(self):
        self.shared.barrier.wait()


class ShardedSample:
    """One rank's view of a frame-sharded inpaint job: its slice of (latent, noise, mask), an engine whose early
    stopper reduces its statistics over the group, and k-diffusion's Euler loop around it (host schedule)."""

    def __init__(self, group, engine, sched, frame_axis: int = 2):
        self.group, self.engine, self.sched, self.axis = group, engine, sched, frame_axis
        engine.stats_reduce = group.all_reduce_sum_

    def my_slice(self, full: torch.Tensor) -> torch.Tensor:
        sl = frame_slice(full.shape[self.axis], self.group.world, self.group.rank)
        idx = [slice(None)] * full.ndim
        idx[self.axis] = sl
        return full[tuple(idx)].contiguous()

    def run(self, latent_image: torch.Tensor, noise: torch.Tensor, latent_mask: torch.Tensor,
            model_options: Optional[dict] = None, tapes: Optional[Sequence] = None) -> torch.Tensor:
        """latent_image / noise / latent_mask are THIS rank's slices.  Returns this rank's slice of the result.
        model_options may carry lanpaint_semantic_stop; every rank takes the same stop decision at every check
        (the statistics are reduced before the threshold test), so the ranks stay in lock step."""
        from .runner import euler_inpaint
        eng = self.engine
        if tapes is not None:
            eng.rng = tapes
        sampling = eng.inner_model.inner_model.model_sampling
        x = sampling.noise_scaling(self.sched.sigmas[0].to(latent_image.device), noise, latent_image)
        for st in self.sched.steps:
            den = eng(x, latent_image, noise, st.sigma_t, latent_mask, st.times, model_options or {}, 0, n_steps=st.n_inner)
            x = torch.add(x, x - den, alpha=(st.sigma_next - st.sigma) / st.sigma)
        return x


class FrameShardedRun:
    """bench.py's record for BASELINE configs[4]: one Wan-sized sample, frames split over the ranks, flow
    simple-20 (shift 3) x N=5, pointwise synthetic denoiser.  With `early_stop_threshold > 0` the stopper runs
    (un-fused half-advances, one reduction kernel + one all_reduce of 2 doubles per check)."""

    def __init__(self, replica_group, dev, latent=(16, 21, 80, 45), n_inner=5, steps=20, shift=3.0,
                 early_stop_threshold: float = 0.0):
        import minicomfy
        from .engine import LanPaint
        from .runner import FlowSampling, HostSchedule, SynthDenoiser
        self.group = DistGroup(replica_group)
        self.dev = dev
        self.latent = tuple(latent)
        sig = minicomfy.simple_scheduler(minicomfy.ModelSamplingCONST(shift), steps)
        self.sched = HostSchedule([float(v) for v in sig], 1, n_inner, flow=True)
        self.threshold = float(early_stop_threshold)
        eng = LanPaint(SynthDenoiser(FlowSampling()), NSteps=n_inner, Friction=15.0, Lambda=5.0, Beta=1.0, StepSize=0.2,
                       IS_FLOW=True, MinStepFrac=1.0, rng="philox", batched_replace="per_sample",
                       EarlyStopThreshold=self.threshold, EarlyStopPatience=1,
                       cuda_graph=True)   # one graph per outer step; with the stopper on the engine launches eagerly
        self.sample = ShardedSample(self.group, eng, self.sched, frame_axis=2)
        g = torch.Generator().manual_seed(5)
        full = (1,) + self.latent
        y = torch.randn(full, generator=g)
        noise = torch.randn(full, generator=g)
        known = (torch.rand((1, 1) + self.latent[1:], generator=g) < 0.5).float()
        self.y, self.noise = self.sample.my_slice(y).to(dev), self.sample.my_slice(noise).to(dev)
        self.known = self.sample.my_slice(known).to(dev)

    def bench(self, jobs: int = 10):
        torch.cuda.default_generators[self.dev.index or 0].manual_seed(100 + self.group.rank)
        eng = self.sample.engine
        for _ in range(2):
            self.sample.run(self.y, self.noise, self.known)
        self.group.barrier()
        eng.reset_counters()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(jobs):
            self.sample.run(self.y, self.noise, self.known)
        e1.record()
        torch.cuda.synchronize(self.dev)
        self.group.barrier()
        ms = self.group.max_over_ranks(e0.elapsed_time(e1))
        done = eng.substeps_done / jobs
        return {"latent": [1] + list(self.latent), "frames_per_rank": int(self.y.shape[2]), "n_gpus": self.group.world,
                "schedule": "flow simple-20 shift 3.0, N=5", "substeps_scheduled": self.sched.substeps,
                "substeps_done_per_job": done, "early_stop_threshold": self.threshold,
                "ms_per_job": ms / jobs, "value": done / (ms / jobs * 1e-3), "unit": "sub-steps/s (one sample)",
                "scaling": "strong (one sample, frames split)",
                "launch": ("one CUDA graph per outer step (engine.cuda_graph)" if self.threshold <= 0 else
                           "plain launches, un-fused half-advances, one host read-back per early-stop check"),
                "collectives": "none" if self.threshold <= 0 else "one all_reduce of 2 doubles per early-stop check"}
