// Host-side coefficient table for the fused Langevin kernels.
//
// Replaces, for one outer diffusion step, the per-sub-step scalar-tensor work of
// the reference: LanPaint.prepare_step_size (src/LanPaint/lanpaint.py:295-328),
// the step-size pin (lanpaint.py:81), the A/D/dt mask blend (lanpaint.py:212-214)
// and the exp/expm1/where/sqrt/clamp of advance_time_overdamped
// (lanpaint.py:241-251).  Because A and dt take exactly one value per
// (sample, mask class), all of it collapses into LP_TABLE_STRIDE floats per row,
// computed here once in fp64 and rounded once to fp32.
#include <cmath>
#include <cstdint>

#include "lanpaint_b200.h"

namespace {

struct Advance {
  double e, k, sd;
};

// One exact Ornstein-Uhlenbeck advance over h for dx = (-A x + C) dt + sqrt(2) dW
// (lanpaint.py:241-251, D = sqrt(2) from lanpaint.py:326-327).
Advance ou_coefficients(double A, double h) {
  Advance a;
  const double Ah = A * h;
  a.e = std::exp(-Ah);
  const bool tiny = std::fabs(A) < 1e-8;
  a.k = tiny ? h : (-std::expm1(-Ah)) / A;
  const double k2 = tiny ? h : (-std::expm1(-2.0 * Ah)) / (2.0 * A);
  const double var = 2.0 * k2;
  a.sd = std::sqrt(var > 0.0 ? var : 0.0);
  return a;
}

void fill_class(float* c, double A, double g, double dt, float* sdm, float* sdmf) {
  const Advance full = ou_coefficients(A, dt);
  const Advance half = ou_coefficients(A, 0.5 * dt);
  // two independent kicks sd1*xi1 then (after multiplication by e_half) sd_half*xi2 == one kick of this std
  *sdm = static_cast<float>(std::sqrt(half.e * half.sd * half.e * half.sd + half.sd * half.sd));
  *sdmf = static_cast<float>(std::sqrt(half.e * full.sd * half.e * full.sd + half.sd * half.sd));
  c[LP_C_G] = static_cast<float>(g);
  c[LP_C_DT] = static_cast<float>(dt);
  c[LP_C_EF] = static_cast<float>(full.e);
  c[LP_C_KF] = static_cast<float>(full.k);
  c[LP_C_SF] = static_cast<float>(full.sd);
  c[LP_C_EH] = static_cast<float>(half.e);
  c[LP_C_KH] = static_cast<float>(half.k);
  c[LP_C_SH] = static_cast<float>(half.sd);
}

}  // namespace

namespace {

// One table row from explicit per-class time steps (shared by both builders).
void fill_row(float* t, double abt, double S, double rep_noise, double rep_y, double corr, double lam,
              double lam_in_target, double dt_free, double dt_known) {
  const double one_m = 1.0 - abt;
  const double inv1m = 1.0 / one_m;
  // lanpaint.py:313-318: A_x = 1/(1-abt), A_y = (1+Lambda)/(1-abt)
  const double A_free = inv1m;
  const double A_known = (1.0 + lam) * inv1m;
  // Coef_C (lanpaint.py:217-220): C = (sqrt(abt) x0e - x_t)/(1-abt) + A x_t = c_tgt * x0e + (A - 1/(1-abt)) * x_t
  t[LP_T_CTGT] = static_cast<float>(std::sqrt(abt) * inv1m);
  t[LP_T_S] = static_cast<float>(S);
  t[LP_T_INVS] = static_cast<float>(1.0 / S);
  t[LP_T_LAM] = static_cast<float>(lam_in_target);
  t[LP_T_ONEPLAM] = static_cast<float>(1.0 + lam_in_target);
  t[LP_T_REPN] = static_cast<float>(rep_noise);
  t[LP_T_REPY] = static_cast<float>(rep_y);
  t[LP_T_CORR] = static_cast<float>(corr);
  fill_class(t + LP_T_CLS0, A_free, 0.0, dt_free, t + LP_T_SDM, t + LP_T_SDMF);
  fill_class(t + LP_T_CLS1, A_known, lam * inv1m, dt_known, t + LP_T_SDM + 1, t + LP_T_SDMF + 1);
  for (int j = 28; j < LP_TABLE_STRIDE; ++j) t[j] = 0.f;
}

}  // namespace

extern "C" int lp_build_coef_table(const double* abt, const double* ve_sigma, const double* rep_noise,
                                   const double* rep_y, const double* corr, int64_t n_rows,
                                   const lp_hyper* hp, float* table_out) {
  if (!abt || !ve_sigma || !hp || !table_out || n_rows < 0) return LP_ERR_INVALID;
  for (int64_t r = 0; r < n_rows; ++r) {
    const double a = abt[r];
    const double one_m = 1.0 - a;
    // lanpaint.py:81  step_size = StepSize * clamp(1 - abt, min=MinStepFrac)
    const double h = hp->step_size * (one_m < hp->min_step_frac ? hp->min_step_frac : one_m);
    // lanpaint.py:301-302,185-190: dtx/2 = h * sigma_x (=1), dty/2 = h * sigma_y (=Beta)
    const double S = hp->flow ? 1.0 / (std::sqrt(a) + std::sqrt(one_m))          // lanpaint.py:97,146,163
                              : std::sqrt(1.0 + ve_sigma[r] * ve_sigma[r]);       // lanpaint.py:99,148,168
    fill_row(table_out + r * LP_TABLE_STRIDE, a, S, rep_noise ? rep_noise[r] : 0.0, rep_y ? rep_y[r] : 1.0,
             corr ? corr[r] : 1.0, hp->lam, hp->lam, h, h * hp->beta);
  }
  return LP_OK;
}

extern "C" int lp_build_coef_table_dt(const double* abt, const double* scale, const double* dt_free,
                                      const double* dt_known, double lam, int32_t target_is_x0e, int64_t n_rows,
                                      float* table_out) {
  if (!abt || !dt_free || !dt_known || !table_out || n_rows < 0) return LP_ERR_INVALID;
  for (int64_t r = 0; r < n_rows; ++r)
    fill_row(table_out + r * LP_TABLE_STRIDE, abt[r], scale ? scale[r] : 1.0, 0.0, 1.0, 1.0, lam,
             target_is_x0e ? 0.0 : lam, dt_free[r], dt_known[r]);
  return LP_OK;
}
