"""Timeline evidence for the host-side gap under CUDA-graph replay (north_star: < 5 %).

nsys is not in the image; torch.profiler (kineto / CUPTI activity records) is.  For each batch size this runs a few
graph-replayed jobs through the node API (`LanPaint_KSampler.sample`, minicomfy standing in for ComfyUI), collects
every GPU kernel record of the sampler loop -- from the first to the last `lp::` kernel of a job -- and reports

    busy  = length of the union of the kernel intervals
    span  = last kernel end - first kernel start
    gap   = 1 - busy / span            (time the GPU sat idle inside the loop: launch gaps, host work between graphs)

    python profiles/timeline_probe.py --requests 1 8 128 [--rng torch] > profiles/r02_timeline.txt
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import minicomfy  # noqa: E402

minicomfy.install()
from lanpaint_b200 import comfy_nodes as N  # noqa: E402
from lanpaint_b200.runner import SynthCondNet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--requests", type=int, nargs="+", default=[1, 8, 128])
ap.add_argument("--rng", default="torch")
ap.add_argument("--jobs", type=int, default=5)
ap.add_argument("--network", default="synth", choices=["synth", "dit"],
                help="synth: the one-kernel pointwise network on SDXL [R,4,128,128]; dit: the bf16 DiT stand-in "
                     "(minicomfy.networks) on the Flux-shaped [R,16,128,128] latent, flow simple-20, cfg 1")
ap.add_argument("--no-callback", action="store_true", help="call the guider without a progress callback: whole-job graph")
ap.add_argument("--per-step", action="store_true",
                help="one graph per outer step with the callback between them even for short jobs (deferred_callbacks=False)")
args = ap.parse_args()
dev = torch.device("cuda:0")


def kernel_records(prof):
    """(name, start_us, end_us) of every GPU kernel record."""
    out = []
    try:
        for e in prof.profiler.kineto_results.events():
            if str(e.device_type()).endswith("CUDA") and e.duration_ns() > 0:
                out.append((e.name(), e.start_ns() / 1e3, (e.start_ns() + e.duration_ns()) / 1e3))
        if out:
            return out
    except Exception:
        pass
    for e in prof.events():
        if str(getattr(e, "device_type", "")).endswith("CUDA"):
            out.append((e.name, e.time_range.start, e.time_range.end))
    return out


def union_length(iv):
    iv = sorted(iv)
    total, cur_s, cur_e = 0.0, None, None
    for s, e in iv:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                total += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        total += cur_e - cur_s
    return total


print(f"# torch.profiler (CUPTI) timeline of the sampler loop, node API, rng={args.rng}, "
      f"{'whole-job graph (no callback)' if args.no_callback else ('one graph per outer step, ComfyUI progress callback between them' if args.per_step else 'default: ComfyUI progress callback; short jobs run as one graph, callbacks delivered after it')}")
print("# requests  job  kernels  span_us  busy_us  gap_%   largest_idle_us")
for R in args.requests:
    g = torch.Generator().manual_seed(0)
    dit = args.network == "dit"
    y = torch.randn(R, 16 if dit else 4, 128, 128, generator=g)
    noise_mask = (torch.rand(R, 1, 128, 128, generator=g) < 0.5).float()
    if dit:
        from minicomfy.networks import DiTStandIn
        torch.manual_seed(1234)
        base = minicomfy.BaseModel(DiTStandIn().to(dev).eval(), model_type=minicomfy.ModelType.FLUX, latent_channels=16,
                                   shift=1.15)
    else:
        base = minicomfy.BaseModel(SynthCondNet())
    patcher = minicomfy.ModelPatcher(base, dev)
    patcher.model_options["lanpaint_b200"] = {"rng": args.rng, "deferred_callbacks": not args.per_step}
    node = N.LanPaint_KSampler()

    def call(seed):
        if not args.no_callback:
            with torch.no_grad():
                return node.sample(patcher, seed, 20, 1.0 if dit else 5.0, "euler", "simple" if dit else "karras", 0.3, -0.2,
                                   {"samples": y, "noise_mask": noise_mask}, 1.0, 5, "Image First", "", N.IMAGE_MODE)
        N._set_hyper(patcher, num_steps=5, cfg=5.0, prompt_mode="Image First")
        guider = minicomfy.CFGGuider(patcher)
        guider.set_conds(0.3, -0.2)
        guider.set_cfg(5.0)
        sig = minicomfy.get_sigmas_karras(20, 0.0292, 14.6146)
        with N.override_sample_function():
            return guider.sample(minicomfy.prepare_noise(y, seed), y, minicomfy.ksampler("euler"), sig,
                                 denoise_mask=noise_mask, seed=seed)
    for k in range(4):      # eager, capture, replay
        call(k)
    torch.cuda.synchronize()
    gaps = []
    for j in range(args.jobs):
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            call(100 + j)
            torch.cuda.synchronize()
        rec = [(n, s, e) for (n, s, e) in kernel_records(prof)]
        lp = [(s, e) for (n, s, e) in rec if "lp::" in n and "pack_mask" not in n]   # the sampler loop proper
        if not lp:
            print(f"{R:9d} {j:4d}  no lp:: kernel records (profiler unavailable?)")
            continue
        t0, t1 = min(s for s, _ in lp), max(e for _, e in lp)
        inside = sorted((s, e) for (n, s, e) in rec if s >= t0 and e <= t1)
        busy, span = union_length(inside), t1 - t0
        idle = max([b[0] - a[1] for a, b in zip(inside, inside[1:])] + [0.0])
        gaps.append(100.0 * (1.0 - busy / span))
        print(f"{R:9d} {j:4d} {len(inside):8d} {span:8.1f} {busy:8.1f} {gaps[-1]:6.2f} {idle:10.2f}   mode={N.LAST_RUN['mode']}")
    if gaps:
        gaps.sort()
        print(f"# requests={R}: median gap {gaps[len(gaps) // 2]:.2f} %")
