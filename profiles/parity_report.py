"""Prints the measured deviation of the CUDA path from the REFERENCE's own outputs (tests/golden/*.npz, generated
by tests/golden/make_golden.py from /root/reference) for every golden case, plus full-size checks against the
oracle on the GPU.  Run on the B200 box:  python profiles/parity_report.py > profiles/r01_parity_report.txt"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from conftest import golden_names, load_golden  # noqa: E402
from _support import make_model, max_rel, rel_l2, synth_inputs  # noqa: E402
from lanpaint_b200.engine import LanPaint, NoiseTape  # noqa: E402
from oracle import langevin_oracle as O  # noqa: E402

dev = torch.device("cuda:0")
print(f"{'case':36s} {'max|d|/max|ref| out':>20s} {'x (in place)':>14s} {'rel-L2 x':>10s}")
worst = 0.0
for name in golden_names():
    g = load_golden(name)
    m = g["meta"]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    x = t(g["x"])
    eng = LanPaint(make_model(m["model"], m["flow"]), m["n_steps"], m["friction"], m["lam"], m["beta"], m["step_size"],
                   IS_FLOW=m["flow"], MinStepFrac=m["min_step_frac"], rng=NoiseTape([t(d) for d in g["tape"]]))
    out = eng(x, t(g["y"]), t(g["noise"]), t(g["sigma"]), t(g["mask_full"]), (t(g["ve"]), t(g["abt"]), t(g["flow_t"])),
              {}, 0, n_steps=m["n_steps"])
    e1, e2 = max_rel(out, torch.from_numpy(g["out"])), max_rel(x, torch.from_numpy(g["x_new"]))
    worst = max(worst, e1, e2)
    print(f"{name:36s} {e1:20.3e} {e2:14.3e} {rel_l2(x, torch.from_numpy(g['x_new'])):10.2e}")
print(f"worst over the reference goldens: {worst:.3e}   (contract: 1e-3)")
print()
print("full-size shapes, same tape, oracle run on the GPU:")
for shape, flow, n in (((8, 4, 128, 128), False, 5), ((32, 4, 128, 128), False, 10), ((1, 16, 128, 128), True, 5),
                       ((1, 16, 21, 80, 45), True, 5)):
    x, y, noise, mk = synth_inputs(shape, seed=11, device=dev)
    sig = torch.full((shape[0],), 0.7 if flow else 2.5, device=dev)
    times = O.times_from_sigma(sig, flow)
    tape = O.NoiseTape(generator=torch.Generator().manual_seed(5))
    out_o, x_o = O.outer_step(make_model("two_heads", flow), x.clone(), y, noise, sig, mk.expand(shape).contiguous(),
                              times, O.Hyper(n_steps=n, min_step_frac=1.0, flow=flow), n_steps=n, draw=tape)
    eng = LanPaint(make_model("two_heads", flow), n, 15.0, 5.0, 1.0, 0.2, IS_FLOW=flow, MinStepFrac=1.0,
                   rng=NoiseTape(tape.recorded))
    xe = x.clone()
    out_e = eng(xe, y, noise, sig, mk, tuple(times), {}, 0, n_steps=n)
    print(f"  {str(shape):24s} N={n:2d} {'flow' if flow else 'VE':4s}  out {max_rel(out_e, out_o):.3e}  x {max_rel(xe, x_o):.3e}  "
          f"rel-L2 {rel_l2(xe, x_o):.2e}")
