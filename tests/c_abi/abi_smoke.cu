// A torch-free, Python-free consumer of the C ABI: plain cudaMalloc'd pointers in, status codes out.
// Runs prologue -> 3 fused sub-steps (tape noise, identity "model": x0 = x0_big = x) -> epilogue and
// compares with a scalar double-precision restatement of the closed form in SURVEY 8a
// (src/LanPaint/lanpaint.py:85-99,182-184,217-220,232-254,274-286).  Exit code 0 = agree.
//
//   nvcc -O2 -Iinclude -o abi_smoke tests/c_abi/abi_smoke.cu -Llanpaint_b200/_lib -llanpaint_b200
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "lanpaint_b200.h"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("cuda error %s at %d\n", cudaGetErrorString(e), __LINE__); return 2; } } while (0)
#define LP(x) do { int rc = (x); if (rc != LP_OK) { printf("lp error %d (%s) at %d\n", rc, lp_status_string(rc), __LINE__); return 3; } } while (0)

static double lcg(unsigned long long& s) {  // uniform (0,1)
  s = s * 6364136223846793005ull + 1442695040888963407ull;
  return ((s >> 11) + 0.5) / 9007199254740992.0;
}
static double gauss(unsigned long long& s) { return std::sqrt(-2.0 * std::log(lcg(s))) * std::cos(6.283185307179586 * lcg(s)); }

int main(int argc, char** argv) {
  if (argc > 1 && std::strcmp(argv[1], "--link-only") == 0) { printf("abi %d\n", lp_abi_version()); return 0; }
  const int B = 2, C = 4, S = 64, per = C * S, n = B * per, N = 3;
  const double sigma[B] = {0.7, 3.0}, lam = 5.0, step = 0.2;
  unsigned long long seed = 12345;
  std::vector<float> x(n), y(n), noise(n);
  std::vector<unsigned char> mask(B * S);
  std::vector<std::vector<float>> tape(2 * N - 1, std::vector<float>(n));
  for (int i = 0; i < n; ++i) { x[i] = (float)gauss(seed); y[i] = (float)gauss(seed); noise[i] = (float)gauss(seed); }
  for (auto& m : mask) m = lcg(seed) < 0.5;
  for (auto& t : tape) for (auto& v : t) v = (float)gauss(seed);

  // ---- host table through the library ----
  double abt[B], ve[B], rn[B], ry[B];
  for (int b = 0; b < B; ++b) { ve[b] = sigma[b]; abt[b] = 1.0 / (1.0 + sigma[b] * sigma[b]); rn[b] = sigma[b]; ry[b] = 1.0; }
  lp_hyper hp = {step, lam, 1.0, 1.0, 0, 0};
  std::vector<float> table(B * LP_TABLE_STRIDE);
  LP(lp_build_coef_table(abt, ve, rn, ry, nullptr, B, &hp, table.data()));
  if (lp_selftest_index_math(100000) != 0) { printf("index math self test failed\n"); return 4; }

  // ---- device side: raw pointers only ----
  float *dx, *dy, *dn, *dc, *dout, *dtab, *dtape;
  unsigned char* dm;
  CK(cudaMalloc(&dx, n * 4)); CK(cudaMalloc(&dy, n * 4)); CK(cudaMalloc(&dn, n * 4)); CK(cudaMalloc(&dc, n * 4));
  CK(cudaMalloc(&dout, n * 4)); CK(cudaMalloc(&dtab, table.size() * 4)); CK(cudaMalloc(&dm, mask.size()));
  CK(cudaMalloc(&dtape, (size_t)tape.size() * n * 4));
  CK(cudaMemcpy(dx, x.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dy, y.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dn, noise.data(), n * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dtab, table.data(), table.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dm, mask.data(), mask.size(), cudaMemcpyHostToDevice));
  for (size_t k = 0; k < tape.size(); ++k) CK(cudaMemcpy(dtape + k * n, tape[k].data(), n * 4, cudaMemcpyHostToDevice));
  cudaStream_t st;
  CK(cudaStreamCreate(&st));
  lp_dims dims = {B, per, S, S, 0, 0};
  LP(lp_prologue_f32(dx, dy, dn, dm, dx, nullptr, dtab, &dims, st));
  int draw = 0;
  for (int i = 0; i < N; ++i) {
    const int flags = (i == 0 ? LP_SUBSTEP_FIRST : 0) | (i + 1 < N ? LP_SUBSTEP_FUSE_NEXT : 0);
    lp_rng r = {LP_RNG_TAPE, 0, dtape + (size_t)draw * n, (i + 1 < N) ? dtape + (size_t)(draw + 1) * n : nullptr, 0, 0, 0, nullptr};
    draw += (i + 1 < N) ? 2 : 1;
    // identity denoiser: both heads are the model input itself (aliased pointers are allowed)
    LP(lp_substep_f32(dx, dx, dx, dy, dm, dc, nullptr, nullptr, dtab, &dims, &r, flags, st));
  }
  LP(lp_epilogue_f32(dx, dy, dm, dout, &dims, st));
  CK(cudaStreamSynchronize(st));
  std::vector<float> gx(n), gout(n);
  CK(cudaMemcpy(gx.data(), dx, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(gout.data(), dout, n * 4, cudaMemcpyDeviceToHost));

  // ---- scalar double restatement ----
  double worst = 0, scale = 0;
  for (int b = 0; b < B; ++b) {
    const double a = abt[b], om = 1 - a, Sx = std::sqrt(1 + sigma[b] * sigma[b]), h = step * (om < 1.0 ? 1.0 : om);
    for (int r = 0; r < per; ++r) {
      const int i = b * per + r, m = mask[b * S + r % S];
      const double A = (m ? 1 + lam : 1.0) / om, dt = h;
      auto adv = [&](double xt, double hh, double Cc, double xi) {
        return std::exp(-A * hh) * xt + (-std::expm1(-A * hh)) / A * Cc + std::sqrt(2 * (-std::expm1(-2 * A * hh)) / (2 * A)) * xi;
      };
      double xm = m ? (double)y[i] + sigma[b] * noise[i] : (double)x[i];
      double xt = xm / Sx, Cp = 0;
      int d = 0;
      for (int k = 0; k < N; ++k) {
        const double x0 = xt * Sx;                                   // identity model in model space
        const double tgt = m ? (1 + lam) * y[i] - lam * x0 : x0;
        const double Cn = (std::sqrt(a) * tgt - xt) / om + A * xt;
        if (k == 0) { xt = adv(xt, dt, Cn, tape[d++][i]); }
        else { xt += (Cn - Cp) * dt; xt = adv(xt, dt / 2, Cp, tape[d++][i]); }
        Cp = Cn;
        if (k + 1 < N) xt = adv(xt, dt / 2, Cn, tape[d++][i]);
      }
      const double xf = xt * Sx, of = m ? (double)y[i] : xf;
      worst = std::fmax(worst, std::fmax(std::fabs(gx[i] - xf), std::fabs(gout[i] - of)));
      scale = std::fmax(scale, std::fabs(xf));
    }
  }
  printf("abi_smoke: max abs err %.3e (max |x| %.3f), abi %d\n", worst, scale, lp_abi_version());
  return worst <= 2e-5 * scale ? 0 : 1;
}
