"""Pin the oracle: (1) against golden vectors produced by the REAL reference engine
(tests/golden/make_golden.py ran /root/reference's LanPaint with a recorded noise
tape), (2) against the known answers the reference's own tests assert
(tests/test_av_schedule.py, tests/test_min_step_frac.py, tests/test_sho_regression.py
in the reference tree), (3) against the schedule facts SURVEY 8d derives from it.
CPU only."""
import pytest
import torch

from conftest import golden_names, load_golden
from _support import make_model
from oracle import langevin_oracle as O


@pytest.mark.parametrize("name", golden_names())
def test_oracle_reproduces_reference_bit_for_bit(name):
    g = load_golden(name)
    meta = g["meta"]
    t = torch.from_numpy
    hp = O.Hyper(n_steps=meta["n_steps"], lam=meta["lam"], beta=meta["beta"], step_size=meta["step_size"],
                 min_step_frac=meta["min_step_frac"], flow=meta["flow"], friction=meta["friction"])
    times = O.Times(t(g["ve"]), t(g["abt"]), t(g["flow_t"]))
    tape = O.NoiseTape([t(d) for d in g["tape"]])
    out, x_new = O.outer_step(make_model(meta["model"], meta["flow"]), t(g["x"]).clone(), t(g["y"]), t(g["noise"]),
                              t(g["sigma"]), t(g["mask_full"]), times, hp, n_steps=meta["n_steps"], draw=tape)
    assert tape.pos == meta["n_draws"]
    assert torch.equal(out, t(g["out"]))
    assert torch.equal(x_new, t(g["x_new"]))


def test_golden_set_covers_the_quirks():
    names = set(golden_names())
    for need in ("cfg1_ve_identity_n5", "ve_two_heads_n0", "ve_zero_noise_regen", "ve_batch2_flowform_replace",
                 "ve_bare_n3", "ve_one_tuple_n3", "flow_two_heads_n5", "video5d_flow_n3", "odd_size_ve_n3",
                 "ve_mask_all_known", "ve_mask_none_known"):
        assert need in names
    # draw order: 1 draw for sub-step 0, 2 for every later one; +1 when the noise image is regenerated
    assert load_golden("ve_two_heads_n5")["meta"]["n_draws"] == 9
    assert load_golden("ve_two_heads_n0")["meta"]["n_draws"] == 0
    assert load_golden("ve_zero_noise_regen")["meta"]["n_draws"] == 1 + 3


def test_oracle_fp64_close_to_fp32():
    """fp32-vs-fp64 drift of the loop itself is ~2e-7 relative."""
    g = load_golden("cfg1_ve_identity_n5")
    t = torch.from_numpy
    hp = O.Hyper(n_steps=5, min_step_frac=1.0)
    times64 = O.Times(*(t(g[k]).double() for k in ("ve", "abt", "flow_t")))
    tape = O.NoiseTape([t(d) for d in g["tape"]])
    out64, x64 = O.outer_step(make_model("identity", False), t(g["x"]).double(), t(g["y"]).double(),
                              t(g["noise"]).double(), t(g["sigma"]).double(), t(g["mask_full"]).double(), times64, hp,
                              n_steps=5, draw=tape)
    rel = float((x64 - t(g["x_new"]).double()).norm() / x64.norm())
    assert rel < 2e-6


# ---- known answers restated from the reference's own tests ---------------------------------
def _flat_pack():
    x = torch.zeros(1, 1, 8)
    ai = torch.zeros(1, 1, 8)
    ai[..., 5:] = 1.0
    video = O.Times(torch.tensor([1.0]), torch.tensor([0.5]), torch.tensor([0.5]))
    audio = O.Times(torch.tensor([0.25]), torch.tensor([0.9]), torch.tensor([0.2]))
    return x, torch.zeros_like(x), torch.ones_like(x), torch.tensor([0.5]), video, audio, ai


def test_audio_rows_get_audio_schedule_parameters():
    """reference tests/test_av_schedule.py:157-199 -> abt 0.5 / 0.9 and step 0.2*(1-abt) per row with the engine's
    default MinStepFrac = 0; without audio context the video schedule applies everywhere."""
    x, y, noise, sigma, video, audio, ai = _flat_pack()
    hp = O.Hyper(n_steps=1, lam=1.0, min_step_frac=0.0)
    _, abt, _, rep, step, _, _ = O.per_row_schedule(sigma, video, hp, O.Audio(ai, audio, None), 3)
    abt, step, rep = abt.flatten(), step.flatten(), rep.flatten()
    assert abt[0] == 0.5 and abt[-1] == pytest.approx(0.9)
    assert step[0] == pytest.approx(0.2 * 0.5) and step[-1] == pytest.approx(0.2 * (1 - 0.9))
    assert rep[0] == pytest.approx(0.5) and rep[-1] == pytest.approx(0.2)
    _, abt2, _, _, step2, _, _ = O.per_row_schedule(sigma, video, hp, None, 3)
    assert abt2.flatten()[0] == 0.5 and abt2.numel() == 1 and step2.flatten()[0] == pytest.approx(0.1)


def test_replace_step_uses_audio_sigma_on_audio_rows():
    """reference tests/test_av_schedule.py:204-219 -> 0.5 on video rows, 0.2 on audio rows."""
    x, y, noise, sigma, video, audio, ai = _flat_pack()
    model = O.IdentityDenoiser(O.FlowSampling())
    hp = O.Hyper(n_steps=0, lam=1.0, flow=False)
    O.outer_step(model, x, y, noise, sigma, torch.ones_like(x), video, hp, n_steps=0,
                 audio=O.Audio(ai, audio, None))
    inp = model.last_input.flatten()
    assert inp[0] == pytest.approx(0.5) and inp[-1] == pytest.approx(0.2)


def test_score_corrects_audio_target_only():
    """reference tests/test_av_schedule.py:244-277 -> 2.0 on video rows, 1.25 on audio rows."""
    class Offset:
        def __init__(self):
            self.inner_model = self
            self.model_sampling = O.FlowSampling()

        def __call__(self, x, sigma, model_options=None, seed=None):
            return x + 2.0, x + 2.0
    ai = torch.zeros(1, 1, 8)
    ai[..., 5:] = 1.0
    corr = (1.0 - ai) + 0.625 * ai
    hp = O.Hyper(n_steps=1, lam=1.0, flow=True)
    z = torch.zeros(1, 1, 8)
    s = O.masked_score(Offset(), z, z, z, torch.full((1, 1, 8), 0.5), torch.ones(1, 1, 8),
                       torch.tensor([0.5]).view(1, 1, 1), hp, 3, correction=corr).flatten()
    assert s[0] == pytest.approx(2.0) and s[-1] == pytest.approx(1.25)
    s2 = O.masked_score(Offset(), z, z, z, torch.full((1, 1, 8), 0.5), torch.ones(1, 1, 8),
                        torch.tensor([0.5]).view(1, 1, 1), hp, 3).flatten()
    assert s2[0] == pytest.approx(2.0) and s2[-1] == pytest.approx(2.0)


def test_step_size_invariant_per_row():
    """reference tests/test_av_schedule.py:304-324 -> A_x * dtx == 0.2 on every row."""
    abt = torch.full((1, 1, 8), 0.5)
    abt[..., 5:] = 0.9
    step = torch.full((1, 1, 8), 0.1)
    step[..., 5:] = 0.02
    one = torch.ones(1, 1, 8)
    cf = O.branch_coefficients(abt, step, one, one, 1.0)
    adt = (cf.A_x * cf.half_dt_x).flatten()  # the tuple slot the reference test calls dtx is dtx/2
    assert adt[0] == pytest.approx(0.2) and adt[-1] == pytest.approx(0.2)


@pytest.mark.parametrize("n,frac,mn,want", [
    (5, 0.1, 0.0, 5), (5, 0.01, 0.0, 5), (5, 0.2, 0.05, 5), (5, 0.05, 0.05, 5), (5, 0.04, 0.05, 4),
    (5, 0.025, 0.05, 2), (5, 0.005, 0.05, 0), (5, 0.0, 0.05, 0), (0, 0.01, 0.05, 0)])
def test_min_step_frac_table(n, frac, mn, want):
    """reference tests/test_min_step_frac.py:17-41 (banker's rounding: 2.5 -> 2, 0.5 -> 0)."""
    assert O.ramped_inner_steps(n, frac, mn) == want


def test_first_order_state_has_no_velocity():
    """reference tests/test_sho_regression.py:6-33."""
    torch.manual_seed(0)
    x = torch.randn(1, 4, 8, 8)
    hp = O.Hyper(n_steps=10, friction=1.0, lam=1.0, beta=1.0, step_size=0.1)
    times = O.Times(torch.tensor([0.5]), torch.tensor([0.5]), torch.tensor([0.5]))
    one = torch.ones(1, 1, 1, 1)
    xo, st = O.langevin_substep(x, lambda z: torch.zeros_like(z), torch.zeros_like(x), torch.tensor([0.1]).view(1, 1, 1, 1),
                                times, hp, None, torch.randn_like, one, one)
    assert st.v is None and st.C is not None and st.x0 is not None
    assert torch.isfinite(xo).all()


# ---- schedule facts -----------------------------------------------------------
def test_sdxl_karras20_substep_counts():
    sig = O.karras_sigmas(20)
    for n, want_seq, want_sub in ((5, [5, 5, 5, 5, 5, 5, 5, 4, 4, 3, 3, 2, 1, 1, 0, 0, 0, 0, 0, 0], 53),
                                  (10, None, 106)):
        seq = []
        for i in range(20):
            s = sig[i] * torch.ones(1)
            tm = O.times_from_sigma(s, False)
            seq.append(O.inner_steps_for(s, sig, tm.abt, n, early_stop=1, min_frac=1.0))
        if want_seq is not None:
            assert seq == want_seq
        assert sum(seq) == want_sub
        assert sum(seq) + 20 == want_sub + 20  # model calls = sub-steps + one final denoise per outer step


def test_euler_inpaint_counts_model_calls():
    sig = O.karras_sigmas(20)
    model = O.PointwiseDenoiser(O.VESampling())
    g = torch.Generator().manual_seed(0)
    y = torch.randn(1, 4, 8, 8, generator=g)
    noise = torch.randn(1, 4, 8, 8, generator=g)
    dm = (torch.rand(1, 1, 8, 8, generator=g) < 0.5).float().expand(1, 4, 8, 8)
    counters = {}
    tape = O.NoiseTape(generator=torch.Generator().manual_seed(1))
    x = O.euler_inpaint(model, y, noise, dm, sig, O.Hyper(n_steps=5, min_step_frac=1.0), draw=tape, counters=counters)
    assert counters == {"substeps": 53, "model_calls": 73}
    assert model.calls == 73 and torch.isfinite(x).all()
