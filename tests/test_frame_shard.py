"""cfg5 frame sharding.

CPU (gloo, world 2): the early stopper's cross-shard reduction -- two ranks each hold half the frames, compute
their two masked sums locally, all_reduce them, and must take exactly the decisions (and report exactly the
distances) of a stopper that sees the whole latent.
GPU (one device, two threads): the whole sharded path -- engines, kernels, stopper -- against the unsharded run
with the same noise tape: bit-equal result, identical stop decisions."""
import os
import socket
import threading

import pytest
import torch
import torch.multiprocessing as mp

from lanpaint_b200.frame_shard import ThreadGroup, frame_slice


def test_frame_slices_cover_the_axis():
    for t in (21, 5, 8):
        for world in (1, 2, 4, 8):
            sl = [frame_slice(t, world, r) for r in range(world)]
            assert sl[0].start == 0 and sl[-1].stop == t
            assert all(sl[i].stop == sl[i + 1].start for i in range(world - 1))
    assert [frame_slice(21, 8, r).stop - frame_slice(21, 8, r).start for r in range(8)] == [3, 3, 3, 3, 3, 2, 2, 2]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cpu_stopper(mask_u8, reduce, trace):
    """An EarlyStopper whose LOCAL sums are computed with torch on the CPU (test double for lp_stop_stats_f32);
    the reduction, weights and decision logic are the product's."""
    from lanpaint_b200 import _native
    from lanpaint_b200.earlystop import EarlyStopper
    counts = torch.tensor([float((mask_u8 == 0).sum()), 0.0], dtype=torch.float64)
    if reduce is not None:
        reduce(counts)
    st = EarlyStopper(threshold=0.05, threshold_eff=0.05, patience_eff=2, mask=mask_u8, ring=None,
                      w_inpaint=float(counts[0]) * 4, w_ring=None, dims=_native.Dims(1, 1, 1, 1, 0, 0), distance_fn=None,
                      trace=trace, bench_ids=(None, None, None), abt_val=0.5, device=torch.device("cpu"), reduce=reduce)

    def local_sums(a, b, table=None):
        free = (mask_u8 == 0).expand(a.shape)
        st._sums[0] = ((a - b).double() ** 2)[free].sum()
        st._sums[1] = 0.0
        return st._sums
    st._local_sums = local_sums
    return st


def _sequence(seed=0, steps=8, shape=(1, 4, 8, 6, 5)):
    g = torch.Generator().manual_seed(seed)
    mask = (torch.rand((1, 1) + shape[2:], generator=g) < 0.5).to(torch.uint8)
    base = torch.randn(shape, generator=g)
    xs = [base + 0.5 * (0.3 ** k) * torch.randn(shape, generator=g) for k in range(steps)]     # converging x0 estimates
    return mask, xs


def _decisions(stopper, xs):
    out, prev = [], None
    for i, cur in enumerate(xs):
        out.append(stopper.step(i=i, n_steps=len(xs), x_before=cur if prev is None else None, x_after=cur,
                                x0_prev=prev, x0_cur=cur, table=None))
        prev = cur
    return out


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from lanpaint_b200.frame_shard import DistGroup
    from lanpaint_b200.replicas import ReplicaGroup
    grp = DistGroup(ReplicaGroup(backend="gloo"))
    mask, xs = _sequence()
    sl = frame_slice(mask.shape[2], world, rank)
    trace = []
    st = _cpu_stopper(mask[:, :, sl].contiguous(), grp.all_reduce_sum_, trace)
    dec = _decisions(st, [x[:, :, sl].contiguous() for x in xs])
    q.put((rank, dec, [t["dist"] for t in trace], st.w_inpaint))
    grp.g.close()


def test_sharded_stopper_takes_the_unsharded_decisions_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    mask, xs = _sequence()
    trace = []
    want = _decisions(_cpu_stopper(mask, None, trace), xs)
    for rank, dec, dists, w in got:
        assert dec == want and any(want), (rank, dec, want)
        assert dists == pytest.approx([t["dist"] for t in trace], rel=1e-12)
        assert w == float((mask == 0).sum()) * 4


@pytest.mark.gpu
@pytest.mark.parametrize("threshold", [0.0, 50.0])
def test_frame_sharded_run_equals_unsharded_run(threshold, cuda_device):
    """Two shards (threads of this process, one GPU) vs the whole sample, same noise tape: bit-equal latents,
    identical early-stop decisions, and -- with the stopper off -- no collective at all."""
    from lanpaint_b200.engine import LanPaint, NoiseTape
    from lanpaint_b200.frame_shard import ShardedSample
    from lanpaint_b200.runner import FlowSampling, HostSchedule, SynthDenoiser
    dev = cuda_device
    shape = (1, 16, 6, 16, 24)
    g = torch.Generator().manual_seed(9)
    y, noise = torch.randn(shape, generator=g).to(dev), torch.randn(shape, generator=g).to(dev)
    known = (torch.rand((1, 1) + shape[2:], generator=g) < 0.5).float().to(dev)
    sig = [1.0, 0.9, 0.75, 0.55, 0.35, 0.15, 0.0]
    sched = HostSchedule(sig, 1, 4, flow=True)
    n_draws = 200
    tape = [torch.randn(shape, generator=g).to(dev) for _ in range(n_draws)]

    def engine():
        return LanPaint(SynthDenoiser(FlowSampling()), NSteps=4, Friction=15.0, Lambda=5.0, Beta=1.0, StepSize=0.2,
                        IS_FLOW=True, MinStepFrac=1.0, batched_replace="per_sample", EarlyStopThreshold=threshold,
                        EarlyStopPatience=1)

    def run(group, out, key):
        torch.cuda.set_device(dev)
        with torch.cuda.stream(torch.cuda.Stream(device=dev)):
            trace = []
            ss = ShardedSample(group, engine(), sched)
            opts = {"lanpaint_semantic_trace": trace}
            res = ss.run(ss.my_slice(y), ss.my_slice(noise), ss.my_slice(known), model_options=opts,
                         tapes=NoiseTape([ss.my_slice(t) for t in tape]))
            torch.cuda.current_stream().synchronize()
            out[key] = (res, [(t["inner_step"], t["stopped"], t["dist"]) for t in trace], ss.engine.substeps_done)

    out = {}
    run(ThreadGroup.make(1)[0], out, "whole")
    groups = ThreadGroup.make(2)
    ths = [threading.Thread(target=run, args=(groups[r], out, r)) for r in range(2)]
    for t in ths:
        t.start()
    for t in ths:
        t.join(timeout=300)
    whole, tr_whole, done_whole = out["whole"]
    assert torch.equal(torch.cat([out[0][0], out[1][0]], dim=2), whole)
    assert out[0][2] == out[1][2] == done_whole
    if threshold > 0:
        assert done_whole < sched.substeps and len(tr_whole) > 0          # the stopper did cut sub-steps
        for r in (0, 1):
            assert [(a, b) for a, b, _ in out[r][1]] == [(a, b) for a, b, _ in tr_whole]
            assert [d for _, _, d in out[r][1]] == pytest.approx([d for _, _, d in tr_whole], rel=1e-6)   # fp32 partial-sum order
    else:
        assert done_whole == sched.substeps
