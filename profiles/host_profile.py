"""cProfile of one graph-replayed node call (host side): where the Python time of the per-sigma wrapper path goes."""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import minicomfy  # noqa: E402

minicomfy.install()
from lanpaint_b200 import comfy_nodes as N  # noqa: E402
from lanpaint_b200.runner import SynthCondNet  # noqa: E402

sampler = sys.argv[1] if len(sys.argv) > 1 else "heun"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
y = torch.randn(R, 4, 128, 128, generator=g)
nm = (torch.rand(R, 1, 128, 128, generator=g) < 0.5).float()
patcher = minicomfy.ModelPatcher(minicomfy.BaseModel(SynthCondNet()), dev)
node = N.LanPaint_KSampler()


def call(seed):
    return node.sample(patcher, seed, 20, 5.0, sampler, "karras", 0.3, -0.2, {"samples": y, "noise_mask": nm}, 1.0, 5,
                       "Image First", "", N.IMAGE_MODE)


for k in range(5):
    call(k)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for k in range(5):
    call(10 + k)
torch.cuda.synchronize()
pr.disable()
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
