"""Bit-equality of a torch-stream kernel geometry (lp_set_option("tma", V)) against the LDG kernels (tma=0)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from lanpaint_b200 import _native  # noqa: E402
from lanpaint_b200.engine import LanPaint  # noqa: E402
from lanpaint_b200.runner import SynthDenoiser, VESampling  # noqa: E402
from lanpaint_b200.schedule import times_from_sigma  # noqa: E402

V = int(sys.argv[1])
dev = torch.device("cuda:0")
lib = _native.load()
ok = True
for shape in ((24, 4, 128, 128), (1, 16, 21, 80, 45), (5, 4, 128, 128), (8, 4, 96, 112)):
    g = torch.Generator().manual_seed(1)
    x, y, nz = (torch.randn(shape, generator=g).to(dev) for _ in range(3))
    m = (torch.rand((shape[0], 1) + shape[2:], generator=g) < 0.5).float().to(dev)
    sig = torch.full((shape[0],), 1.3)
    res = []
    for tma in (0, V):
        lib.lp_set_option(b"tma", tma)
        torch.manual_seed(5)
        e = LanPaint(SynthDenoiser(VESampling()), 4, 15.0, 5.0, 1.0, 0.2, MinStepFrac=1.0, rng="torch", batched_replace="per_sample")
        xx = x.clone()
        res.append((e(xx, y, nz, sig, m, times_from_sigma(sig, False), None, 0, n_steps=4), xx))
    same = torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    ok &= same
    print(shape, "bit-identical" if same else "MISMATCH")
print("variant", V, "OK" if ok else "FAILED")
