"""Node-API probe: K calls of LanPaint_KSampler.sample (minicomfy standing in for ComfyUI) on a batch of R SDXL
requests, against runner.GraphedJob on the same guider.  Prints wall time per call, the device time of the
sampler loop inside the call, and where the rest of the wall time goes."""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import minicomfy  # noqa: E402

minicomfy.install()
from lanpaint_b200 import comfy_nodes as N  # noqa: E402
from lanpaint_b200.runner import SynthCondNet  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--requests", type=int, default=128)
ap.add_argument("--rng", default="torch")
ap.add_argument("--calls", type=int, default=8)
ap.add_argument("--sampler", default="euler")
ap.add_argument("--no-graph", action="store_true")
ap.add_argument("--no-fused", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
R = args.requests
g = torch.Generator().manual_seed(0)
y = torch.randn(R, 4, 128, 128, generator=g)
noise_mask = (torch.rand(R, 1, 128, 128, generator=g) < 0.5).float()
patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(SynthCondNet()), dev)
patcher.model_options["lanpaint_b200"] = {"rng": args.rng, "timing": True, "cuda_graph": not args.no_graph,
                                          "fused_sampler": not args.no_fused}
node = N.LanPaint_KSampler()
for k in range(args.calls):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    (out,) = node.sample(patcher, 100 + k, 20, 5.0, args.sampler, "karras", 0.3, -0.2,
                         {"samples": y, "noise_mask": noise_mask}, 1.0, 5, "Image First", "", N.IMAGE_MODE)
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3
    e0, e1 = N.LAST_RUN["events"]
    print(f"call {k}: mode={N.LAST_RUN['mode']} fused={N.LAST_RUN['fused']} wall {wall:.2f} ms, "
          f"sampler loop on device {e0.elapsed_time(e1):.3f} ms, "
          f"model calls {N.LAST_ENGINE['engine'].model_calls}, launches {N.LAST_ENGINE['engine'].launches}")
t0 = time.perf_counter()
minicomfy.prepare_noise(y, 1)
print(f"prepare_noise (CPU randn, ComfyUI's own): {(time.perf_counter() - t0) * 1e3:.2f} ms")
