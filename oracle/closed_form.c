/* TEST INFRASTRUCTURE -- a second, independent CPU oracle in plain C (double precision, scalar loops).
 *
 * Restates ONE outer step of LanPaint's inner loop as the per-element closed form of SURVEY.md 8a, i.e. the
 * arithmetic the CUDA kernels implement, without any of the reference's tensor-op structure:
 *   replace step + VP map        src/LanPaint/lanpaint.py:85-99
 *   masked score target          src/LanPaint/lanpaint.py:182-184  (x_t + score = x0 | (1+lam) y - lam x0_BIG)
 *   Coef_C                       src/LanPaint/lanpaint.py:217-220
 *   exact OU advance             src/LanPaint/lanpaint.py:232-254
 *   first / steady sub-step      src/LanPaint/lanpaint.py:274-286  (old C in the second half-advance)
 *   epilogue                     src/LanPaint/lanpaint.py:144-157
 * with the pointwise two-head stand-in denoiser of the tests (h0 = a0 x + b0 tanh x + c0, h1 = a1 x + c1).
 * tests/test_oracle_closed_form.py checks it against oracle/langevin_oracle.py (which is pinned bit-for-bit to
 * the reference) in fp64: two restatements written differently must agree to round-off.
 * Only tests/ may load this.  Build: make -C oracle   ->  oracle/_build/libclosed_form.so
 */
#include <math.h>
#include <stdint.h>

typedef struct {
  double step_size, lam, beta, min_step_frac;
  int32_t flow;      /* IS_FLUX or IS_FLOW */
  int32_t n_steps;
  double coef[5];    /* a0 b0 c0 a1 c1 */
} cf_params;

static double ou(double x, double h, double A, double C, double xi) {
  const double e = exp(-A * h);
  const double k = fabs(A) < 1e-8 ? h : -expm1(-A * h) / A;
  const double k2 = fabs(A) < 1e-8 ? h : -expm1(-2.0 * A * h) / (2.0 * A);
  const double var = 2.0 * k2; /* D = sqrt(2) */
  return e * x + k * C + sqrt(var > 0.0 ? var : 0.0) * xi;
}

/* x, y, noise: [B][per] ; mask: [B][per] (1 = known); sigma, abt, ve: [B]; rep_noise/rep_y: [B] replace form;
 * tape: [n_draws][B][per], consumed 1 draw for sub-step 0 and 2 for each later one.
 * Outputs: out [B][per] (returned tensor), x_new [B][per] (the in-place rewritten x).  Returns draws used. */
int cf_outer_step(const double* x, const double* y, const double* noise, const uint8_t* mask, const double* abt,
                  const double* ve, const double* rep_noise, const double* rep_y, const double* tape, int64_t B,
                  int64_t per, const cf_params* p, double* out, double* x_new) {
  const int64_t n = B * per;
  int draws = 0;
  for (int64_t b = 0; b < B; ++b) {
    const double a = abt[b], om = 1.0 - a;
    const double S = p->flow ? 1.0 / (sqrt(a) + sqrt(om)) : sqrt(1.0 + ve[b] * ve[b]);
    const double h = p->step_size * (om < p->min_step_frac ? p->min_step_frac : om);
    for (int64_t r = 0; r < per; ++r) {
      const int64_t i = b * per + r;
      const int m = mask[i] != 0;
      const double A = (m ? 1.0 + p->lam : 1.0) / om;
      const double dt = m ? h * p->beta : h;
      double xm = m ? rep_noise[b] * noise[i] + rep_y[b] * y[i] : x[i];
      double xt, Cp;
      int d;
      /* ---- the loop in the reference's own order: half-step, model, correct, half-step with OLD C ---- */
      xt = xm / S;
      Cp = 0.0;
      d = 0;
      for (int k = 0; k < p->n_steps; ++k) {
        if (k > 0) xt = ou(xt, 0.5 * dt, A, Cp, tape[(int64_t)(d++) * n + i]);
        const double xin = xt * S;
        const double h0 = p->coef[0] * xin + p->coef[1] * tanh(xin) + p->coef[2];
        const double h1 = p->coef[3] * xin + p->coef[4];
        const double tgt = m ? (1.0 + p->lam) * y[i] - p->lam * h1 : h0;
        const double Cn = (sqrt(a) * tgt - xt) / om + A * xt;
        if (k == 0) {
          xt = ou(xt, dt, A, Cn, tape[(int64_t)(d++) * n + i]);
        } else {
          xt += (Cn - Cp) * dt;
          xt = ou(xt, 0.5 * dt, A, Cp, tape[(int64_t)(d++) * n + i]);
        }
        Cp = Cn;
      }
      const double xf = xt * S;
      const double o0 = p->coef[0] * xf + p->coef[1] * tanh(xf) + p->coef[2];
      x_new[i] = xf;
      out[i] = m ? y[i] : o0;
      draws = d;
    }
  }
  return draws;
}
