#!/usr/bin/env python
"""bench.py -- Langevin sub-steps/sec on the SDXL 128x128x4 latent (BASELINE.json's metric).

    python bench.py --gpus N --steps K --warmup W              # this repo (CUDA kernels)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU path (oracle port)

Workload (`config.workload`): `requests_per_gpu` independent SDXL inpaint requests of
shape [1,4,128,128] (BASELINE configs[1]'s latent), batched per GPU, each running the
reference schedule: karras-20 sigmas x N=5 think steps with the node defaults
(MinStepFrac=1, EarlyStop=1) = 53 Langevin sub-steps + 20 final denoises = 73 model
calls per request, with the SURVEY-8d synthetic pointwise two-head denoiser
standing in for the UNet and k-diffusion's Euler update between outer steps.
One bench "step" = one such job over the GPU's whole batch.  value = request-sub-steps/s
summed over GPUs (weak scaling: per-GPU batch fixed).

Timed with CUDA events on the launching stream, barrier + synchronize on both sides, max
over ranks.  Inputs of the default batch (128 requests: 176 MB touched per launch) exceed
the 126 MB L2, so every launch streams from HBM (`config.l2`).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

SHAPE = (4, 128, 128)          # SDXL latent of a 1024x1024 image
N_OUTER, N_INNER = 20, 5
ALGO_BYTES_PER_ELEM = 28.0 + 1.0 / SHAPE[0]   # read x,x0,x0B,y,C + write x,C (fp32) + uint8 spatial mask
METRIC = "Langevin sub-steps/sec (SDXL 128x128x4 latent, N=5)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            time.sleep(0.15)
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.1)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = max(mx, float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        load = [v for v in sm if v > 0.5 * mx] or sm
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(index: int):
    """Best effort: run this rank (and first-touch its pinned buffers) on the NUMA node its GPU hangs off, so
    eight ranks do not push their H2D/D2H traffic through one socket.  Returns the node or None."""
    try:
        p = torch.cuda.get_device_properties(index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if allowed:
            os.sched_setaffinity(0, allowed)
            return node
    except Exception:
        pass
    return None


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------
# the reference arm: the oracle (op-for-op restatement of the reference's eager PyTorch path)
# on the host cores.  The one place outside tests/ where oracle/ is executed.
# ------------------------------------------------------------------------------------------
def calibrate_threads(requests: int) -> int:
    """The reference's eager path is ~90 small element-wise ops per sub-step; on a many-core host more
    threads can be much slower (fork/join per op).  Time ONE outer step of the actual workload (5 sub-steps,
    6 model calls) per candidate thread count and keep the fastest, so the CPU arm is the best this host can
    do rather than a strawman."""
    cores = os.cpu_count() or 1
    best, best_t = 1, float("inf")
    for c in sorted({k for k in (2, 4, 8, 16, 32, 64, cores) if k <= cores}):
        warm = cpu_job(requests, 1, c)[0]      # also warms this thread count
        if warm > 2.5 * best_t:                # past the knee: more threads only get slower on this host
            break
        dt = min(cpu_job(requests, 1, c)[0] for _ in range(2))
        if dt < best_t:
            best, best_t = c, dt
    return best


def cpu_job(requests: int, outer_steps: int, threads: int, seed: int = 0):
    """Runs the first `outer_steps` outer steps of the workload on `requests` requests on the CPU.
    Returns (seconds, request-sub-steps done)."""
    from oracle import langevin_oracle as O
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(seed)
    shape = (requests,) + SHAPE
    y, noise = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    dm = (torch.rand((requests, 1) + SHAPE[1:], generator=g) < 0.5).float().expand(shape).contiguous()
    sig = O.karras_sigmas(N_OUTER)
    hp = O.Hyper(n_steps=N_INNER, min_step_frac=1.0)
    model = O.PointwiseDenoiser(O.VESampling())
    counters = {}
    t0 = time.perf_counter()
    _cpu_partial(O, model, y, noise, dm, sig, hp, min(outer_steps, N_OUTER), counters)
    dt = time.perf_counter() - t0
    return dt, counters["substeps"] * requests


def _cpu_partial(O, model, y, noise, dm, sig, hp, outer_steps, counters):
    """First `outer_steps` outer steps with the FULL schedule's step bookkeeping (n_eff depends on
    the position in the full 20-step schedule, nodes.py:286-299)."""
    x = model.model_sampling.noise_scaling(sig[0], noise, y)
    mask = O.binarise_mask(dm)
    s_in = x.new_ones([x.shape[0]])
    sub = 0
    for i in range(outer_steps):
        sigma = sig[i] * s_in
        tm = O.times_from_sigma(sigma, False)
        n_eff = O.inner_steps_for(sigma, sig, tm.abt, hp.n_steps, 1, 1.0)
        den, x = O.outer_step(model, x, y, noise, sigma, mask, tm, hp, n_eff)
        x = x + (x - den) / sigma.view(-1, 1, 1, 1) * (sig[i + 1] - sig[i])
        sub += n_eff
    counters["substeps"] = sub


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    host_cores = os.cpu_count() or 1
    req = args.ref_requests
    cores = calibrate_threads(req)
    for _ in range(args.warmup):
        cpu_job(req, N_OUTER, cores)
    t, units = 0.0, 0
    for _ in range(args.steps):
        dt, u = cpu_job(req, N_OUTER, cores)
        t += dt
        units += u
    value = units / t
    sample = (f"{req} of {args.requests} requests per step, full karras-20 x N=5 schedule (53 sub-steps/request); "
              f"{cores} torch threads (fastest of a probe over 1..{host_cores} host cores)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "sub-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args),
        "cpu_baseline": {"value": value, "unit": "sub-steps/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "sub-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args):
    touched = 5.25 * 4 * args.requests * SHAPE[0] * SHAPE[1] * SHAPE[2] / 1e6
    return {"workload": "sdxl_1024_inpaint_4x128x128_karras20_N5", "requests_per_gpu": args.requests,
            "latent_shape": [1] + list(SHAPE), "outer_steps": N_OUTER, "think_steps": N_INNER,
            "substeps_per_request": 53, "model_calls_per_request": 73, "denoiser": "synthetic pointwise two-head",
            "sampler": "euler",
            "mask": "random 50% per spatial site" if args.mask == "random" else "centred 90x90 hole (49.4% unknown)",
            "rng": args.rng,
            "launch": {"job-graph": "one CUDA graph per job (20 outer steps), replayed per request batch",
                       "step-graph": "one CUDA graph per outer step", "eager": "plain launches"}[args.launch],
            "parallelism": f"replicas x{args.gpus} (requests sharded, no data-path collective)",
            "l2": (f"inputs larger than L2, no flush: {touched:.0f} MB touched per sub-step launch vs 126 MB L2 (launches "
                   "stream mostly from HBM; config.sweep shows the L2-resident sizes and the fully HBM-bound R=256; "
                   "the roofline probe cycles 3 operand sets so it is L2-cold)")
                  if touched > 126 else f"working set {touched:.0f} MB fits L2; no flush (see config.sweep for HBM-bound size)"}


# ------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------
MASK_KIND = "random"


def make_inputs(requests, dev, seed, pinned=False):
    g = torch.Generator().manual_seed(seed)
    shape = (requests,) + SHAPE
    y = torch.randn(shape, generator=g)
    noise = torch.randn(shape, generator=g)
    if MASK_KIND == "blob":   # a real inpainting mask: known everywhere except one centred 90x90 hole (49.4 %)
        mask = torch.ones((requests, 1) + SHAPE[1:])
        mask[:, :, 19:109, 19:109] = 0.0
    else:                     # rand(B,1,H,W) < 0.5 per spatial site (worst case for operand skipping)
        mask = (torch.rand((requests, 1) + SHAPE[1:], generator=g) < 0.5).float()  # 1 = known
    if pinned:
        return [t.pin_memory() for t in (y, noise, mask)]
    return [t.to(dev) for t in (y, noise, mask)]


def run_b200(args):
    import torch.distributed as dist
    from lanpaint_b200.engine import LanPaint, pack_mask
    from lanpaint_b200.runner import GraphedJob, HostSchedule, SynthDenoiser, VESampling, euler_inpaint, karras_sigmas

    rank, world, local = dist_env()
    assert torch.cuda.is_available(), "bench.py (impl b200) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_numa_node(local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    R = args.requests
    model = SynthDenoiser(VESampling())
    # replicas: the only collective is the one-time broadcast of the denoiser's weights (north_star)
    w = torch.tensor(model.coef, device=dev)
    if world > 1:
        dist.broadcast(w, src=0)
    model.set_coef(w.tolist())

    def make_engine(graph):
        return LanPaint(model, NSteps=N_INNER, Friction=15.0, Lambda=5.0, Beta=1.0, StepSize=0.2, MinStepFrac=1.0,
                        rng=args.rng, batched_replace="per_sample", cuda_graph=graph)
    eng = make_engine(args.launch == "step-graph")
    sched = HostSchedule(karras_sigmas(N_OUTER), R, N_INNER)
    assert sched.substeps == 53 and sched.model_calls == 73
    y, noise, mask = make_inputs(R, dev, seed=rank)
    pm = pack_mask(mask, y)
    torch.manual_seed(1000 + rank)

    gjob = GraphedJob(eng, sched, (R,) + SHAPE, dev, l2_persist=args.l2_persist) if args.launch == "job-graph" else None

    def job():
        if gjob is not None:
            return gjob.run(y, noise, pm)
        return euler_inpaint(eng, y, noise, pm, sched)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        job()
    barrier()

    # ---- timed region: K jobs, device time, max over ranks ----
    eng.launches = 0
    eng.model_calls = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        e0.record()
        for _ in range(args.steps):
            job()
        e1.record()
        barrier()
    ms = e0.elapsed_time(e1)
    tmax = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    ms = float(tmax.item())
    launches = eng.launches + eng.model_calls   # every synthetic-denoiser call is one kernel of this repo
    units = world * R * sched.substeps * args.steps
    value = units / (ms * 1e-3)

    # ---- roofline of the dominant kernel (steady fused sub-step), live CUDA-event timing ----
    peak, peak_src = peaks()
    n_el = R * SHAPE[0] * SHAPE[1] * SHAPE[2]
    algo = ALGO_BYTES_PER_ELEM * n_el
    roof = {"bound": "hbm", "achieved": None, "peak": peak, "unit": "GB/s", "frac": None, "traffic": None,
            "kernel": ("lp::substep_tma_kernel<first=0, merge=1> (steady fused sub-step, TMA-staged persistent)"
                       if args.rng == "philox" and n_el >= (1 << 20) else
                       "lp::substep_kernel<VEC=4, %s, first=0, fuse_next=1> (steady fused sub-step)" % args.rng),
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": algo}
    if args.kernel_timer:
        from lanpaint_b200.runner import time_steady_substep
        # (1) 53 back-to-back launches of the steady kernel on job-shaped operands, x20 (the roofline number)
        burst = sorted(time_steady_substep(eng, y, pm, sigma=2.0, launches=53, repeats=20, rotate=3))
        avg = sum(burst) / len(burst)
        roof.update(achieved=algo / (avg * 1e-6) / 1e9, avg_us=avg, median_us=burst[len(burst) // 2],
                    min_us=burst[0], launches_timed=53 * len(burst),
                    timing="53 back-to-back launches of the steady fused sub-step between two CUDA events on the "
                           "launching stream, x20, cycling 3 independent job-shaped operand sets so every launch's "
                           "operands were evicted from L2 by the two launches before it (true HBM streaming)")
        roof["frac"] = roof["achieved"] / peak
        warm = sorted(time_steady_substep(eng, y, pm, sigma=2.0, launches=53, repeats=10, rotate=1))
        roof["same_buffers_us"] = sum(warm) / len(warm)   # one operand set re-used: the L2 keeps part of it
        # (2) the same kernel inside real jobs: one CUDA-event pair around every launch of an eager pass
        # (includes the ~launch latency an isolated launch pays; reported for the share-of-step cross-check)
        eng_t = make_engine(False)
        for _ in range(2):
            euler_inpaint(eng_t, y, noise, pm, sched)
        barrier()
        eng_t.kernel_timer = timer = []
        for _ in range(max(3, min(10, args.steps))):
            euler_inpaint(eng_t, y, noise, pm, sched)
        barrier()
        eng_t.kernel_timer = None
        mid = sorted(a.elapsed_time(b) * 1e3 for f, a, b in timer if (f & 3) == 2)  # steady = FUSE_NEXT, not FIRST
        if mid:
            roof["in_job_event_pair_us"] = sum(mid) / len(mid)
        roof["substep_share_of_step"] = 53 * avg * 1e-3 / (ms / args.steps)
        tr = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tr):
            roof["traffic"] = json.load(open(tr)).get(str(R))

    if args.kernel_timer and rank == 0:
        # context for the fraction above: the same copy probe MEASURED_PEAKS.json was produced with
        # (torch b.copy_(a) over 1 Gi bf16 elements, read+write bytes, best of 10), on THIS box
        try:
            a_ = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev)
            b_ = torch.empty_like(a_)
            best = float("inf")
            for _ in range(10):
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record()
                b_.copy_(a_)
                c1.record()
                c1.synchronize()
                best = min(best, c0.elapsed_time(c1))
            roof["copy_gbs_this_box"] = 2 * a_.numel() * 2 / (best * 1e-3) / 1e9
            del a_, b_
        except Exception:
            pass

    # ---- e2e: host buffers in, host result out, through the public call; copies inside the timed region.
    # Two request batches are in flight on two streams (each with its own pinned buffers and job graph), so
    # the H2D of batch k+1 and the D2H of batch k-1 overlap the kernels of batch k, as a serving loop does.
    e2e = None
    if not args.no_e2e:
        lanes = []
        for lane in range(2 if gjob is not None else 1):
            hy, hn, hm = make_inputs(R, dev, seed=rank + 17 * lane, pinned=True)
            lanes.append({"in": (hy, hn, hm), "out": torch.empty((R,) + SHAPE).pin_memory(),
                          "stream": torch.cuda.Stream(device=dev),
                          "job": GraphedJob(make_engine(False), sched, (R,) + SHAPE, dev,
                                            l2_persist=args.l2_persist) if gjob is not None else None})
        bi = sum(t.numel() * 4 for t in lanes[0]["in"])
        bo = lanes[0]["out"].numel() * 4

        def e2e_submit(ln):
            hy, hn, hm = ln["in"]
            with torch.cuda.stream(ln["stream"]):
                if ln["job"] is not None:   # pinned host tensors straight into the job's static buffers
                    ln["job"].run(hy, hn, hm.to(dev, non_blocking=True), x_out=ln["out"])
                else:
                    dy, dn, dm = (t.to(dev, non_blocking=True) for t in (hy, hn, hm))
                    euler_inpaint(eng, dy, dn, dm, sched, x_out=ln["out"])

        def e2e_run(n):
            for i in range(n):
                ln = lanes[i % len(lanes)]
                ln["stream"].synchronize()      # the previous result of this lane is on the host: the user has it
                e2e_submit(ln)
            for ln in lanes:
                ln["stream"].synchronize()

        e2e_run(2 * len(lanes))
        barrier()
        k2 = max(4, args.steps // 2)
        t0 = time.perf_counter()
        e2e_run(k2)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": world * R * sched.substeps * k2 / float(tt.item()), "unit": "sub-steps/s",
               "h2d_bytes_per_step": bi, "d2h_bytes_per_step": bo, "steps": k2, "in_flight": len(lanes),
               "numa_node_rank0": numa,
               "api": ("lanpaint_b200.runner.GraphedJob.run" if gjob is not None else "lanpaint_b200.runner.euler_inpaint")
                      + "(engine=lanpaint_b200.LanPaint) on pinned host tensors, result to pinned host memory"}

    # ---- CPU baseline: the oracle port on the host cores, bounded sample (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        host_cores = os.cpu_count() or 1
        cores = calibrate_threads(args.ref_requests)
        cpu_job(2, 2, cores)  # warm
        dt, u = cpu_job(args.ref_requests, N_OUTER, cores)
        # the same port with device="cuda": what a user of the reference runs today on this very GPU
        # (eager PyTorch, ~89 element-wise launches per sub-step).  Informational, not the reference arm.
        gpu_eager = None
        try:
            from oracle import langevin_oracle as O
            rq = args.ref_requests
            yq, nq, mq = make_inputs(rq, dev, seed=3)
            sq = O.karras_sigmas(N_OUTER).to(dev)
            hq = O.Hyper(n_steps=N_INNER, min_step_frac=1.0)
            dmq = (1 - mq).expand_as(yq).contiguous()
            cnt = {}
            O.euler_inpaint(O.PointwiseDenoiser(O.VESampling()), yq, nq, dmq, sq, hq, counters=cnt)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            O.euler_inpaint(O.PointwiseDenoiser(O.VESampling()), yq, nq, dmq, sq, hq, counters=cnt)
            torch.cuda.synchronize()
            tq = time.perf_counter() - t0
            gpu_eager = {"value": rq * cnt["substeps"] / tq, "unit": "sub-steps/s", "requests": rq,
                         "ms_per_job": 1e3 * tq, "what": "oracle port, device=cuda, eager launches, same schedule"}
        except Exception as e:  # never let the informational leg break the bench line
            gpu_eager = {"error": f"{type(e).__name__}: {e}"}
        cpu = {"value": u / dt, "unit": "sub-steps/s", "cores": cores, "kind": "port", "same_port_on_this_gpu": gpu_eager,
               "sample": f"{args.ref_requests} of {R} requests, full karras-20 x N=5 schedule, {dt:.1f} s of CPU work; "
                         f"{cores} torch threads (fastest of a probe over 1..{host_cores} host cores)"}

    # ---- the literal BASELINE configs (batch 1, batch 8) and the L2-resident regime, same method as `value` ----
    sweep = None
    if rank == 0 and world == 1 and args.sweep:
        sweep = []
        for r in (1, 8, 32, 64, 256):
            if r == R:
                continue
            m_r = SynthDenoiser(VESampling())
            e_r = LanPaint(m_r, NSteps=N_INNER, Friction=15.0, Lambda=5.0, Beta=1.0, StepSize=0.2, MinStepFrac=1.0,
                           rng=args.rng, batched_replace="per_sample")
            s_r = HostSchedule(karras_sigmas(N_OUTER), r, N_INNER)
            y_r, n_r, k_r = make_inputs(r, dev, seed=5)
            j_r = GraphedJob(e_r, s_r, (r,) + SHAPE, dev)
            p_r = pack_mask(k_r, y_r)
            for _ in range(3):
                j_r.run(y_r, n_r, p_r)
            torch.cuda.synchronize()
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(40 if r <= 64 else 12):
                j_r.run(y_r, n_r, p_r)
            a1.record()
            torch.cuda.synchronize()
            t_r = a0.elapsed_time(a1) / (40 if r <= 64 else 12)
            sweep.append({"requests_per_gpu": r, "ms_per_step": t_r, "value": r * s_r.substeps / (t_r * 1e-3),
                          "note": "working set fits L2" if r <= 64 else "every launch streams from HBM"})
            del j_r, e_r, m_r, y_r, n_r, k_r, p_r
            torch.cuda.empty_cache()

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "sub-steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args), "roofline": roof, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches * world, "clocks": clocks.summary(),
        }
        if sweep is not None:
            line["config"]["sweep"] = sweep
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--requests", type=int, default=128, help="independent SDXL requests batched per GPU")
    ap.add_argument("--ref-requests", type=int, default=8, help="requests per step in the CPU arm's bounded sample")
    ap.add_argument("--rng", default="philox", choices=["philox", "torch"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-sweep", dest="sweep", action="store_false", help="skip the R=1/8/32 lines in config.sweep")
    ap.add_argument("--no-kernel-timer", dest="kernel_timer", action="store_false")
    ap.add_argument("--l2-persist", action="store_true", help="pin the clean latent in L2 (measured slower; off)")
    ap.add_argument("--launch", default="job-graph", choices=["job-graph", "step-graph", "eager"],
                    help="one CUDA graph per job (default) | one per outer step | plain launches")
    ap.add_argument("--mask", default="random", choices=["random", "blob"],
                    help="random 50%% per site (SURVEY 8d, default) | one centred hole of the same area")
    args = ap.parse_args()
    global MASK_KIND
    MASK_KIND = args.mask
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
