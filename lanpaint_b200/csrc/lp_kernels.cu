// lanpaint_b200: sm_100a kernels for LanPaint's inner Langevin loop + their C ABI.
//
// One fused launch per Langevin sub-step replaces the ~89 element-wise ATen
// kernels the reference issues between two model calls
// (src/LanPaint/lanpaint.py:113-142,159-184,192-293).  The work is a pure
// HBM stream: 28+1/C bytes per latent element, a dozen FMAs, two Gaussian
// draws.  So the design rules are the streaming ones: 128-bit coalesced
// accesses, one table row of host-precomputed coefficients per sample instead
// of per-element exp/expm1/sqrt, Philox + Box-Muller in registers, no shared
// memory round trip, no host synchronisation, graph-capturable.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "lanpaint_b200.h"

namespace lp {

thread_local int g_last_cuda_error = 0;

constexpr int kBlock = 256;

// Programmatic dependent launch: every kernel below starts with pdl_prologue() -- wait until the grid it
// depends on has completed and flushed (griddepcontrol.wait), then let the NEXT kernel of the stream begin
// launching (griddepcontrol.launch_dependents) -- and every launch goes through launch_kernel(), which sets
// cudaLaunchAttributeProgrammaticStreamSerialization.  Inside a captured job (146 small dependent kernels)
// this hides most of the launch latency between consecutive nodes; it never relaxes ordering, because each
// dependent still waits for its predecessor to finish before touching memory.
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// process-wide switches, initialised from the environment, adjustable through lp_set_option()
static int g_opt_pdl = [] {
  const char* e = getenv("LANPAINT_B200_PDL");
  return (e && e[0] == '0') ? 0 : 1;
}();
static int g_opt_tma = [] {
  const char* e = getenv("LANPAINT_B200_TMA");
  return e ? atoi(e) : 1;
}();

inline bool pdl_enabled() { return g_opt_pdl != 0; }

template <typename... KArgs, typename... Args>
inline void launch_kernel_smem(void (*kernel)(KArgs...), dim3 grid, size_t smem, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kBlock);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

template <typename... KArgs, typename... Args>
inline void launch_kernel(void (*kernel)(KArgs...), dim3 grid, cudaStream_t stream, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(kBlock);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ----------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011; same constants/round structure as cuRAND's
// curand_philox4x32_x.h so LP_RNG_TORCH can reproduce torch.randn bit for bit)
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox_round(uint4 c, uint2 k) {
  const uint32_t hi0 = __umulhi(0xD2511F53u, c.x);
  const uint32_t lo0 = 0xD2511F53u * c.x;
  const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z);
  const uint32_t lo1 = 0xCD9E8D57u * c.z;
  return make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
}

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 9; ++r) {
    c = philox_round(c, k);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return philox_round(c, k);
}

// Box-Muller with exactly cuRAND's arithmetic (curand_normal.h:_curand_box_muller):
// precise logf / sqrtf, fast __sincosf.  Used by LP_RNG_TORCH.
__device__ __forceinline__ float2 box_muller_curand(uint32_t a, uint32_t b) {
  const float u = a * 2.3283064e-10f + (2.3283064e-10f / 2);
  const float v = b * (2.3283064e-10f * 6.2831855f) + ((2.3283064e-10f * 6.2831855f) / 2);
  const float s = sqrtf(-2.0f * logf(u));
  float2 r;
  __sincosf(v, &r.x, &r.y);
  r.x *= s;
  r.y *= s;
  return r;
}

// Cheaper variant for LP_RNG_PHILOX: MUFU lg2 / sqrt / sin / cos only.
__device__ __forceinline__ float2 box_muller_fast(uint32_t a, uint32_t b) {
  const float u = a * 2.3283064e-10f + (2.3283064e-10f / 2);  // (0, 1]
  const float v = b * (2.3283064e-10f * 6.2831855f) + ((2.3283064e-10f * 6.2831855f) / 2);
  float s;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(s) : "f"(-2.0f * __logf(u)));
  float2 r;
  __sincosf(v, &r.x, &r.y);
  r.x *= s;
  r.y *= s;
  return r;
}

template <bool kCurandExact>
__device__ __forceinline__ float4 normal4(uint4 r) {
  const float2 p = kCurandExact ? box_muller_curand(r.x, r.y) : box_muller_fast(r.x, r.y);
  const float2 q = kCurandExact ? box_muller_curand(r.z, r.w) : box_muller_fast(r.z, r.w);
  return make_float4(p.x, p.y, q.x, q.y);
}

// LP_RNG_PHILOX: element i <- component (i & 3) of Philox(counter = {i>>2, draw}, key = seed).
__device__ __forceinline__ float4 philox_normal4(uint64_t seed, uint64_t draw, uint32_t vec_index) {
  const uint4 c = make_uint4(vec_index, 0u, (uint32_t)draw, (uint32_t)(draw >> 32));
  const uint2 k = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  return normal4<false>(philox4x32_10(c, k));
}

// LP_RNG_TORCH: what curand_init(seed, subsequence, offset) + the (call)-th curand_normal4 yields
// when offset is a multiple of 4 (it always is for torch's generator).
__device__ __forceinline__ float4 torch_normal4(uint64_t seed, uint64_t offset, uint32_t subsequence,
                                                uint32_t call) {
  const uint64_t ctr = (offset >> 2) + call;
  const uint4 c = make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), subsequence, 0u);
  const uint2 k = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
  return normal4<true>(philox4x32_10(c, k));
}

__device__ __forceinline__ float pick(const float4& v, uint32_t j) {
  return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w));
}

// ----------------------------------------------------------------------------
// per-row coefficients (see LP_T_* in lanpaint_b200.h)
// ----------------------------------------------------------------------------
template <bool kFirst, bool kNext>
struct RowCoef {
  float c_tgt, S, inv_S, lam, one_plus_lam, corr;
  float g[2], dt[2], e1[2], k1[2], s1[2];  // advance #1: full dt when kFirst, else half
  float e2[2], k2[2], s2[2];               // advance #2 (only when kNext): always half
  float sm[2];                             // merged kick of advance #1 + #2 (LP_SUBSTEP_MERGE_NOISE)

  // The row is one 128-byte line: eight 128-bit read-only loads, L1-resident after the first warp.
  __device__ __forceinline__ void load(const float* __restrict__ t) {
    const float4* q = reinterpret_cast<const float4*>(t);
    const float4 h0 = __ldg(q + 0), h1 = __ldg(q + 1);
    const float4 a0 = __ldg(q + 2), a1 = __ldg(q + 3), b0 = __ldg(q + 4), b1 = __ldg(q + 5);
    const float4 m = __ldg(q + 6);
    c_tgt = h0.x; S = h0.y; inv_S = h0.z; lam = h0.w;
    one_plus_lam = h1.x; corr = h1.w;
    g[0] = a0.x; dt[0] = a0.y; g[1] = b0.x; dt[1] = b0.y;
    // class row: g, dt, e_full, k_full, sd_full, e_half, k_half, sd_half
    e1[0] = kFirst ? a0.z : a1.y; k1[0] = kFirst ? a0.w : a1.z; s1[0] = kFirst ? a1.x : a1.w;
    e1[1] = kFirst ? b0.z : b1.y; k1[1] = kFirst ? b0.w : b1.z; s1[1] = kFirst ? b1.x : b1.w;
    e2[0] = a1.y; k2[0] = a1.z; s2[0] = a1.w;
    e2[1] = b1.y; k2[1] = b1.z; s2[1] = b1.w;
    sm[0] = kFirst ? m.z : m.x;
    sm[1] = kFirst ? m.w : m.y;
  }
};

// The whole per-element update between two model calls.
//   x      model-space state (in: what the model just saw; out: what it sees next)
//   cprev  drift constant C of the previous sub-step (ignored when kFirst)
//   returns the new C through cnew and x_t + score through x0e
// Reference: score_model lanpaint.py:182-184, Coef_C :217-220,
// advance_time_overdamped :232-254, run_overdamped :274-286.
template <bool kFirst, bool kNext, bool kMerge = false>
__device__ __forceinline__ void substep_element(float& x, float x0, float x0b, float y, float cprev,
                                                bool known, float xi1, float xi2,
                                                const RowCoef<kFirst, kNext>& t, float& cnew,
                                                float& x0e) {
  // per-class coefficients by select (constant indices keep the row in registers)
  const float g = known ? t.g[1] : t.g[0];
  const float dt = known ? t.dt[1] : t.dt[0];
  const float e1 = known ? t.e1[1] : t.e1[0];
  const float k1 = known ? t.k1[1] : t.k1[0];
  const float s1 = known ? t.s1[1] : t.s1[0];
  const float e2 = known ? t.e2[1] : t.e2[0];
  const float k2 = known ? t.k2[1] : t.k2[0];
  const float s2 = known ? t.s2[1] : t.s2[0];
  if (t.corr != 1.0f) {  // audio rows only (lanpaint.py:173-180); uniform per row
    x0 = fmaf(t.corr, x0 - x, x);
    x0b = fmaf(t.corr, x0b - x, x);
  }
  float xt = x * t.inv_S;
  // x_t + score: free region -> x0 ; known region -> (1+lam) y - lam x0_BIG
  const float tgt = known ? fmaf(-t.lam, x0b, t.one_plus_lam * y) : x0;
  const float cn = fmaf(t.c_tgt, tgt, g * xt);
  if (kMerge) {
    // both kicks folded into one Gaussian of std sm (see LP_SUBSTEP_MERGE_NOISE); drift terms unchanged
    const float sm = known ? t.sm[1] : t.sm[0];
    if (kFirst) {
      xt = fmaf(e1, xt, k1 * cn);
    } else {
      xt = fmaf(cn - cprev, dt, xt);
      xt = fmaf(e1, xt, k1 * cprev);
    }
    xt = fmaf(e2, xt, fmaf(k2, cn, sm * xi1));
  } else {
    if (kFirst) {
      xt = fmaf(e1, xt, fmaf(k1, cn, s1 * xi1));
    } else {
      xt = fmaf(cn - cprev, dt, xt);
      xt = fmaf(e1, xt, fmaf(k1, cprev, s1 * xi1));  // old C on purpose (lanpaint.py:283-284)
    }
    if (kNext) xt = fmaf(e2, xt, fmaf(k2, cn, s2 * xi2));
  }
  x = xt * t.S;
  cnew = cn;
  x0e = tgt;
}

// n / d for n < 2^31 without a hardware divide: q = umulhi(n, mul) >> shift (mul == 0: d == 1).
struct FastDiv {
  uint32_t d, mul, shift;
  __device__ __forceinline__ uint32_t div(uint32_t n) const { return mul ? (__umulhi(n, mul) >> shift) : n; }
};

struct Geometry {
  uint32_t total;     // B * per_row
  uint32_t mask_row_stride;
  uint32_t mask_channel_stride;
  FastDiv per_row;    // elements per table row
  FastDiv spatial;    // elements per channel
  uint32_t n_rows;
  uint32_t row_split; // 0 = off; positions >= row_split of a row use table row n_rows + row
};

// element index -> (table row, mask element)
__device__ __forceinline__ void locate(const Geometry& g, uint32_t i, uint32_t& row, uint32_t& mask_index) {
  row = g.per_row.div(i);
  const uint32_t r = i - row * g.per_row.d;
  const uint32_t ch = g.spatial.div(r);
  const uint32_t s = r - ch * g.spatial.d;
  mask_index = row * g.mask_row_stride + ch * g.mask_channel_stride + s;
  if (g.row_split && r >= g.row_split) row += g.n_rows;
}

struct SubstepArgs {
  float* x;
  const float* x0;
  const float* x0b;
  const float* y;
  const uint8_t* mask;
  float* c;
  float* x_copy;
  float* x0e;
  const float* table;
  const float* tape0;
  const float* tape1;
  const uint64_t* rng_state;
  uint64_t seed, draw0, draw1;
  Geometry g;
  uint32_t torch_T;  // threads of torch's randn grid (LP_RNG_TORCH)
  int store_c;       // write C even when the next half-advance is not fused
  int use_cfg;       // x0/x0b hold raw cond/uncond predictions: combine them here
  float cfg, cfg_big;
};

template <int N>
struct Vec;
template <>
struct Vec<1> {
  using F = float;
  using M = uint8_t;
};
template <>
struct Vec<4> {
  using F = float4;
  using M = uchar4;
};

template <int N>
__device__ __forceinline__ void load_f(const float* p, uint32_t i, float (&v)[N]) {
  if (N == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p + i);
    v[0] = t.x; v[1 % N] = t.y; v[2 % N] = t.z; v[3 % N] = t.w;
  } else {
    v[0] = p[i];
  }
}
template <int N>
__device__ __forceinline__ void load_f_ro(const float* __restrict__ p, uint32_t i, float (&v)[N]) {
  if (N == 4) {
    float4 t;  // read-once stream: keep it out of L1
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w)
                 : "l"(p + i));
    v[0] = t.x; v[1 % N] = t.y; v[2 % N] = t.z; v[3 % N] = t.w;
  } else {
    v[0] = __ldg(p + i);
  }
}
template <int N>
__device__ __forceinline__ void store_f(float* p, uint32_t i, const float (&v)[N]) {
  if (N == 4) {
    *reinterpret_cast<float4*>(p + i) = make_float4(v[0], v[1 % N], v[2 % N], v[3 % N]);
  } else {
    p[i] = v[0];
  }
}
template <int N>
__device__ __forceinline__ void load_m(const uint8_t* __restrict__ p, uint32_t i, bool (&v)[N]) {
  if (N == 4) {
    const uchar4 t = __ldg(reinterpret_cast<const uchar4*>(p + i));
    v[0] = t.x != 0; v[1 % N] = t.y != 0; v[2 % N] = t.z != 0; v[3 % N] = t.w != 0;
  } else {
    v[0] = __ldg(p + i) != 0;
  }
}

// ---- the memory side of a fused launch: one N-wide vector per thread ------------
template <int N, bool kFirst, bool kNext, bool kMerge>
__device__ __forceinline__ void substep_vector(const SubstepArgs& a, uint32_t i, const float (&xi1)[N],
                                               const float (&xi2)[N]) {
  uint32_t row, mi;
  locate(a.g, i, row, mi);
  float x[N], x0[N], x0b[N], y[N], cp[N], cn[N], te[N];
  bool known[N];
  // The mask decides which operands this vector needs: free positions use only x0, known positions only
  // x0_big and y (score_model, lanpaint.py:182-184).  Inpainting masks are spatially coherent, so most
  // vectors are all-free or all-known and skip 4-8 of their 20 input bytes per element.
  load_m<N>(a.mask, mi, known);
  load_f<N>(a.x, i, x);
  if (!kFirst) {
    load_f<N>(a.c, i, cp);
  } else {
#pragma unroll
    for (int j = 0; j < N; ++j) cp[j] = 0.f;
  }
  bool any_known = false, any_free = false;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    any_known |= known[j];
    any_free |= !known[j];
  }
  const bool aliased = a.x0b == a.x0;
  const bool need_x0 = any_free || aliased || a.use_cfg;
  const bool need_x0b = (any_known || a.use_cfg) && !aliased;
#pragma unroll
  for (int j = 0; j < N; ++j) x0[j] = x0b[j] = y[j] = 0.f;
  if (need_x0) load_f_ro<N>(a.x0, i, x0);
  if (need_x0b) {
    load_f_ro<N>(a.x0b, i, x0b);
  } else if (aliased) {
#pragma unroll
    for (int j = 0; j < N; ++j) x0b[j] = x0[j];
  }
  if (a.use_cfg) {  // x0 = cond, x0b = uncond: u + (c-u)*s with the eager path's three roundings
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float u = x0b[j], d = __fsub_rn(x0[j], u);
      x0[j] = __fadd_rn(u, __fmul_rn(d, a.cfg));
      x0b[j] = __fadd_rn(u, __fmul_rn(d, a.cfg_big));
    }
  }
  if (any_known) load_f_ro<N>(a.y, i, y);
  RowCoef<kFirst, kNext> t;
  t.load(a.table + (size_t)row * LP_TABLE_STRIDE);
#pragma unroll
  for (int j = 0; j < N; ++j)
    substep_element<kFirst, kNext, kMerge>(x[j], x0[j], x0b[j], y[j], cp[j], known[j], xi1[j], xi2[j], t, cn[j],
                                           te[j]);
  store_f<N>(a.x, i, x);
  if (kNext || a.store_c) store_f<N>(a.c, i, cn);
  if (a.x_copy) store_f<N>(a.x_copy, i, x);
  if (a.x0e) store_f<N>(a.x0e, i, te);
}

// ---- TAPE / PHILOX ------------------------------------------------------------
template <int N, int kRng, bool kFirst, bool kNext, bool kMerge = false>
__global__ void __launch_bounds__(kBlock, 5) substep_kernel(const SubstepArgs a) {
  pdl_prologue();
  const uint32_t v = blockIdx.x * kBlock + threadIdx.x;
  const uint32_t i = v * N;
  if (i >= a.g.total) return;
  float xi1[N], xi2[N];
  if (kRng == LP_RNG_TAPE) {
    load_f_ro<N>(a.tape0, i, xi1);
    if (kNext) {
      load_f_ro<N>(a.tape1, i, xi2);
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) xi2[j] = 0.f;
    }
  } else {
    uint64_t seed = a.seed, d0 = a.draw0, d1 = a.draw1;
    if (a.rng_state) {
      seed = a.rng_state[0];
      d0 += a.rng_state[1];
      d1 += a.rng_state[1];
    }
    const float4 n1 = philox_normal4(seed, d0, i >> 2);
    float4 n2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kNext && !kMerge) n2 = philox_normal4(seed, d1, i >> 2);
    if (N == 4) {
      xi1[0] = n1.x; xi1[1 % N] = n1.y; xi1[2 % N] = n1.z; xi1[3 % N] = n1.w;
      xi2[0] = n2.x; xi2[1 % N] = n2.y; xi2[2 % N] = n2.z; xi2[3 % N] = n2.w;
    } else {
      xi1[0] = pick(n1, i & 3);
      xi2[0] = pick(n2, i & 3);
    }
  }
  substep_vector<N, kFirst, kNext, kMerge>(a, i, xi1, xi2);
}

// ---- PHILOX, TMA-staged persistent variant of the steady fused sub-step ---------------------------
// Same arithmetic and the same Philox stream as substep_kernel (results are bit-identical); different data
// movement: a persistent grid (2 CTAs per SM) walks tiles of kTile consecutive elements of one (row, channel),
// one elected thread streams each tile's five operand slices + mask slice into shared memory with
// cp.async.bulk (the TMA unit, completion counted on an mbarrier), kStages tiles ahead, and the 256 threads
// consume them with LDS.128 and write x / C back with STG.128.  Registers hold no loads in flight, so
// bytes-in-flight per SM is set by kStages * 42 KB instead of by occupancy.
template <int kTile>
struct __align__(128) TmaStage {
  float x[kTile], x0[kTile], x0b[kTile], y[kTile], c[kTile];
  uint8_t m[kTile];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra.uni DONE;\n"
      "bra.uni LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

struct TileGeom {
  uint32_t tiles_per_channel, n_tiles, channels;
};

template <bool kFirst, bool kNext, bool kMerge, int kTile, int kStages, int kMinBlocks>
__global__ void __launch_bounds__(kBlock, kMinBlocks) substep_tma_kernel(const SubstepArgs a, const TileGeom tg) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  using Stage = TmaStage<kTile>;
  Stage* stage = reinterpret_cast<Stage*>(smem_raw);
  __shared__ __align__(8) uint64_t full[kStages];
  pdl_prologue();
  const uint32_t S = a.g.spatial.d;
  const bool aliased = a.x0b == a.x0;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  auto tile_origin = [&](uint32_t tile, uint32_t& e0, uint32_t& len, uint32_t& row, uint32_t& mi0) {
    const uint32_t rc = tile / tg.tiles_per_channel;          // row * channels + channel
    const uint32_t s0 = (tile - rc * tg.tiles_per_channel) * kTile;
    len = S - s0 < (uint32_t)kTile ? S - s0 : (uint32_t)kTile;
    e0 = rc * S + s0;
    row = rc / tg.channels;
    mi0 = row * a.g.mask_row_stride + (rc - row * tg.channels) * a.g.mask_channel_stride + s0;
  };
  auto issue = [&](uint32_t tile, int s) {  // one thread
    uint32_t e0, len, row, mi0;
    tile_origin(tile, e0, len, row, mi0);
    const uint32_t fb = len * 4u;
    const uint32_t total = fb * (3u + (aliased ? 0u : 1u) + (kFirst ? 0u : 1u)) + len;
    mbar_expect_tx(&full[s], total);
    Stage& t = stage[s];
    tma_load_1d(t.x, a.x + e0, fb, &full[s]);
    tma_load_1d(t.x0, a.x0 + e0, fb, &full[s]);
    if (!aliased) tma_load_1d(t.x0b, a.x0b + e0, fb, &full[s]);
    tma_load_1d(t.y, a.y + e0, fb, &full[s]);
    if (!kFirst) tma_load_1d(t.c, a.c + e0, fb, &full[s]);
    tma_load_1d(t.m, a.mask + mi0, len, &full[s]);
  };

  uint64_t seed = a.seed, d0 = a.draw0, d1 = a.draw1;
  if (a.rng_state) {
    seed = a.rng_state[0];
    d0 += a.rng_state[1];
    d1 += a.rng_state[1];
  }
  if (threadIdx.x == 0) {
#pragma unroll
    for (int s = 0; s < kStages; ++s) {
      const uint32_t tile = blockIdx.x + (uint32_t)s * gridDim.x;
      if (tile < tg.n_tiles) issue(tile, s);
    }
  }
  uint32_t k = 0;
  for (uint32_t tile = blockIdx.x; tile < tg.n_tiles; tile += gridDim.x, ++k) {
    const int s = (int)(k % kStages);
    mbar_wait(&full[s], (k / kStages) & 1u);
    uint32_t e0, len, row, mi0;
    tile_origin(tile, e0, len, row, mi0);
    const Stage& t = stage[s];
    RowCoef<kFirst, kNext> rc;
    rc.load(a.table + (size_t)row * LP_TABLE_STRIDE);
#pragma unroll
    for (int pass = 0; pass < kTile / (4 * kBlock); ++pass) {
      const uint32_t v = threadIdx.x + pass * kBlock;  // vector inside the tile
      if (4 * v < len) {
        const float4 xv = reinterpret_cast<const float4*>(t.x)[v];
        float4 x0v = reinterpret_cast<const float4*>(t.x0)[v];
        float4 x0bv = aliased ? x0v : reinterpret_cast<const float4*>(t.x0b)[v];
        const float4 yv = reinterpret_cast<const float4*>(t.y)[v];
        const float4 cv = kFirst ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<const float4*>(t.c)[v];
        const uchar4 mv = reinterpret_cast<const uchar4*>(t.m)[v];
        float x[4] = {xv.x, xv.y, xv.z, xv.w}, x0[4] = {x0v.x, x0v.y, x0v.z, x0v.w};
        float x0b[4] = {x0bv.x, x0bv.y, x0bv.z, x0bv.w}, y[4] = {yv.x, yv.y, yv.z, yv.w};
        float cp[4] = {cv.x, cv.y, cv.z, cv.w}, cn[4], te[4];
        const bool known[4] = {mv.x != 0, mv.y != 0, mv.z != 0, mv.w != 0};
        if (a.use_cfg) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float u = x0b[j], d = __fsub_rn(x0[j], u);
            x0[j] = __fadd_rn(u, __fmul_rn(d, a.cfg));
            x0b[j] = __fadd_rn(u, __fmul_rn(d, a.cfg_big));
          }
        }
        const uint32_t i = e0 + 4 * v;
        const float4 n1 = philox_normal4(seed, d0, i >> 2);
        float4 n2 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (kNext && !kMerge) n2 = philox_normal4(seed, d1, i >> 2);
        const float xi1[4] = {n1.x, n1.y, n1.z, n1.w}, xi2[4] = {n2.x, n2.y, n2.z, n2.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          substep_element<kFirst, kNext, kMerge>(x[j], x0[j], x0b[j], y[j], cp[j], known[j], xi1[j], xi2[j], rc, cn[j],
                                                te[j]);
        *reinterpret_cast<float4*>(a.x + i) = make_float4(x[0], x[1], x[2], x[3]);
        if (kNext || a.store_c) *reinterpret_cast<float4*>(a.c + i) = make_float4(cn[0], cn[1], cn[2], cn[3]);
      }
    }
    __syncthreads();  // every thread is done reading stage s
    const uint32_t next = tile + (uint32_t)kStages * gridDim.x;
    if (threadIdx.x == 0 && next < tg.n_tiles) issue(next, s);
  }
}

// ---- TORCH, 128-bit path --------------------------------------------------------
// torch.randn_like gives element li the component ((li div T) mod 4) of the ((li div 4T))-th
// curand_normal4 of Philox subsequence (li mod T).  Thread (t, k) of this kernel IS torch's thread t at
// its k-th call: it generates that call's four normals (planes 4k..4k+3), then the four lanes of a quad
// transpose them with shuffles so that lane q ends up with plane 4k+q of torch threads t0..t0+3 -- four
// CONSECUTIVE elements, i.e. one float4 -- and runs the same vector body as the other modes.
// One Philox call per 4 elements per draw, 128-bit accesses, bit-identical stream.
// 4x4 transpose across the 4 lanes of a quad, two butterfly stages, register indices all static:
// in: v[j] = my normal for plane j;  out: v[i] = plane q's normal of quad lane i.
__device__ __forceinline__ void quad_transpose(float (&v)[4], uint32_t q) {
  const bool b0 = (q & 1u) != 0, b1 = (q & 2u) != 0;
  // stage 1: 2x2 blocks between lanes q and q^1
  float s0 = b0 ? v[0] : v[1], s1 = b0 ? v[2] : v[3];
  s0 = __shfl_xor_sync(0xffffffffu, s0, 1);
  s1 = __shfl_xor_sync(0xffffffffu, s1, 1);
  v[0] = b0 ? s0 : v[0]; v[1] = b0 ? v[1] : s0;
  v[2] = b0 ? s1 : v[2]; v[3] = b0 ? v[3] : s1;
  // stage 2: 2x2 blocks of pairs between lanes q and q^2
  float u0 = b1 ? v[0] : v[2], u1 = b1 ? v[1] : v[3];
  u0 = __shfl_xor_sync(0xffffffffu, u0, 2);
  u1 = __shfl_xor_sync(0xffffffffu, u1, 2);
  v[0] = b1 ? u0 : v[0]; v[2] = b1 ? v[2] : u0;
  v[1] = b1 ? u1 : v[1]; v[3] = b1 ? v[3] : u1;
}

template <bool kFirst, bool kNext>
__global__ void __launch_bounds__(kBlock) substep_torchvec_kernel(const SubstepArgs a) {
  pdl_prologue();
  const uint32_t T = a.torch_T;
  const uint32_t t = blockIdx.x * kBlock + threadIdx.x;  // torch thread == Philox subsequence (grid.x*256 == T)
  const uint32_t k = blockIdx.y;                          // index of the curand_normal4 call
  const uint32_t q = t & 3u;
  uint64_t seed = a.seed, o0 = a.draw0, o1 = a.draw1;
  if (a.rng_state) {
    seed = a.rng_state[0];
    o0 += a.rng_state[1];
    o1 += a.rng_state[1];
  }
  const float4 n1 = torch_normal4(seed, o0, t, k);
  float4 n2 = make_float4(0.f, 0.f, 0.f, 0.f);
  if (kNext) n2 = torch_normal4(seed, o1, t, k);
  float xi1[4] = {n1.x, n1.y, n1.z, n1.w};
  float xi2[4] = {n2.x, n2.y, n2.z, n2.w};
  quad_transpose(xi1, q);
  if (kNext) quad_transpose(xi2, q);
  const uint64_t e = (uint64_t)(t - q) + (uint64_t)(4u * k + q) * T;  // first of my 4 consecutive elements
  if (e >= a.g.total) return;
  substep_vector<4, kFirst, kNext, false>(a, (uint32_t)e, xi1, xi2);
}

// ---- TORCH, scalar fallback: torch.randn_like's own thread<->element mapping ---
// thread t of T handles elements t, t+T, t+2T, t+3T (one curand_normal4) per
// 4T-stride iteration, exactly like distribution_elementwise_grid_stride_kernel
// (ATen/native/cuda/DistributionTemplates.h), so one Philox call feeds 4
// elements and the stream matches the eager reference on the same generator.
template <bool kFirst, bool kNext>
__global__ void __launch_bounds__(kBlock) substep_torch_kernel(const SubstepArgs a) {
  pdl_prologue();
  const uint32_t T = a.torch_T;
  const uint32_t tid = blockIdx.x * kBlock + threadIdx.x;
  if (tid >= T) return;
  uint64_t seed = a.seed, o0 = a.draw0, o1 = a.draw1;
  if (a.rng_state) {
    seed = a.rng_state[0];
    o0 += a.rng_state[1];
    o1 += a.rng_state[1];
  }
  uint32_t call = 0;
  for (uint64_t base = tid; base < a.g.total; base += 4ull * T, ++call) {
    const float4 n1 = torch_normal4(seed, o0, tid, call);
    float4 n2 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (kNext) n2 = torch_normal4(seed, o1, tid, call);
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const uint64_t li = base + (uint64_t)ii * T;
      if (li >= a.g.total) break;
      const uint32_t i = (uint32_t)li;
      uint32_t row, mi;
      locate(a.g, i, row, mi);
      RowCoef<kFirst, kNext> t;
      t.load(a.table + (size_t)row * LP_TABLE_STRIDE);
      float x = a.x[i];
      float x0 = __ldg(a.x0 + i);
      float x0b = __ldg(a.x0b + i);
      if (a.use_cfg) {
        const float u = x0b, d = __fsub_rn(x0, u);
        x0 = __fadd_rn(u, __fmul_rn(d, a.cfg));
        x0b = __fadd_rn(u, __fmul_rn(d, a.cfg_big));
      }
      const float y = __ldg(a.y + i);
      const bool known = __ldg(a.mask + mi) != 0;
      const float cp = kFirst ? 0.f : a.c[i];
      float cn, te;
      substep_element<kFirst, kNext>(x, x0, x0b, y, cp, known, pick(n1, ii), pick(n2, ii), t, cn, te);
      a.x[i] = x;
      if (kNext || a.store_c) a.c[i] = cn;
      if (a.x_copy) a.x_copy[i] = x;
      if (a.x0e) a.x0e[i] = te;
    }
  }
}

// ---- un-fused OU advance (advance_time_overdamped, lanpaint.py:232-254) --------
struct AdvanceArgs {
  float* x;
  const float* c;
  const uint8_t* mask;
  const float* table;
  const float* tape0;
  const uint64_t* rng_state;
  uint64_t seed, draw0;
  Geometry g;
  uint32_t torch_T;
  int half;
};

__device__ __forceinline__ float advance_element(float x, float c, bool known, float xi,
                                                 const float* __restrict__ t, int half) {
  const float* k = t + (known ? LP_T_CLS1 : LP_T_CLS0);
  const float e = __ldg(k + (half ? LP_C_EH : LP_C_EF));
  const float kk = __ldg(k + (half ? LP_C_KH : LP_C_KF));
  const float sd = __ldg(k + (half ? LP_C_SH : LP_C_SF));
  const float xt = x * __ldg(t + LP_T_INVS);
  return fmaf(e, xt, fmaf(kk, c, sd * xi)) * __ldg(t + LP_T_S);
}

template <int kRng>
__global__ void __launch_bounds__(kBlock) advance_kernel(const AdvanceArgs a) {
  pdl_prologue();
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= a.g.total) return;
  uint32_t row, mi;
  locate(a.g, i, row, mi);
  float xi;
  if (kRng == LP_RNG_TAPE) {
    xi = __ldg(a.tape0 + i);
  } else {
    uint64_t seed = a.seed, d0 = a.draw0;
    if (a.rng_state) {
      seed = a.rng_state[0];
      d0 += a.rng_state[1];
    }
    xi = pick(philox_normal4(seed, d0, i >> 2), i & 3);
  }
  a.x[i] = advance_element(a.x[i], __ldg(a.c + i), __ldg(a.mask + mi) != 0, xi,
                           a.table + (size_t)row * LP_TABLE_STRIDE, a.half);
}

__global__ void __launch_bounds__(kBlock) advance_torch_kernel(const AdvanceArgs a) {
  pdl_prologue();
  const uint32_t T = a.torch_T;
  const uint32_t tid = blockIdx.x * kBlock + threadIdx.x;
  if (tid >= T) return;
  uint64_t seed = a.seed, o0 = a.draw0;
  if (a.rng_state) {
    seed = a.rng_state[0];
    o0 += a.rng_state[1];
  }
  uint32_t call = 0;
  for (uint64_t base = tid; base < a.g.total; base += 4ull * T, ++call) {
    const float4 n1 = torch_normal4(seed, o0, tid, call);
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const uint64_t li = base + (uint64_t)ii * T;
      if (li >= a.g.total) break;
      const uint32_t i = (uint32_t)li;
      uint32_t row, mi;
      locate(a.g, i, row, mi);
      a.x[i] = advance_element(a.x[i], __ldg(a.c + i), __ldg(a.mask + mi) != 0, pick(n1, ii),
                               a.table + (size_t)row * LP_TABLE_STRIDE, a.half);
    }
  }
}

// ---- early-stop statistics: two masked sums of squared differences ---------------
// Grid-stride over float4 groups; per-thread partials -> warp shuffle -> one smem slot per warp ->
// one atomicAdd(double) per block per sum (earlystop.py:51-55 _weighted_mse numerators).
struct StatsArgs {
  const float* a;
  const float* b;
  const uint8_t* mask;
  const uint8_t* ring;
  const float* table;
  double* sums;
  Geometry g;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <int N>
__global__ void __launch_bounds__(kBlock) stop_stats_kernel(const StatsArgs s) {
  pdl_prologue();
  float acc_in = 0.f, acc_ring = 0.f;
  const uint32_t stride = gridDim.x * kBlock * N;
  for (uint32_t i = (blockIdx.x * kBlock + threadIdx.x) * N; i < s.g.total; i += stride) {
    uint32_t row, mi;
    locate(s.g, i, row, mi);
    float av[N], bv[N];
    bool known[N], on_ring[N];
    load_f_ro<N>(s.a, i, av);
    load_f_ro<N>(s.b, i, bv);
    load_m<N>(s.mask, mi, known);
    if (s.ring) {
      load_m<N>(s.ring, mi, on_ring);
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) on_ring[j] = false;
    }
    const float scale = s.table ? __ldg(s.table + (size_t)row * LP_TABLE_STRIDE + LP_T_INVS) : 1.f;
#pragma unroll
    for (int j = 0; j < N; ++j) {
      const float d = (av[j] - bv[j]) * scale;
      const float d2 = d * d;
      acc_in += known[j] ? 0.f : d2;
      acc_ring += on_ring[j] ? d2 : 0.f;
    }
  }
  __shared__ float part[2][kBlock / 32];
  acc_in = warp_sum(acc_in);
  acc_ring = warp_sum(acc_ring);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) {
    part[0][w] = acc_in;
    part[1][w] = acc_ring;
  }
  __syncthreads();
  if (w == 0) {
    float u = lane < kBlock / 32 ? part[0][lane] : 0.f;
    float v = lane < kBlock / 32 ? part[1][lane] : 0.f;
    u = warp_sum(u);
    v = warp_sum(v);
    if (lane == 0) {
      atomicAdd(s.sums + 0, (double)u);
      atomicAdd(s.sums + 1, (double)v);
    }
  }
}

// ---- prologue / epilogue / utilities ----------------------------------------
template <int N>
__global__ void __launch_bounds__(kBlock) prologue_kernel(const float* x, const float* __restrict__ y,
                                                          const float* __restrict__ noise,
                                                          const uint8_t* __restrict__ mask, float* x_model,
                                                          float* x_copy, const float* __restrict__ table,
                                                          Geometry g) {
  pdl_prologue();
  const uint32_t i = (blockIdx.x * kBlock + threadIdx.x) * N;
  if (i >= g.total) return;
  uint32_t row, mi;
  locate(g, i, row, mi);
  const float rn = __ldg(table + (size_t)row * LP_TABLE_STRIDE + LP_T_REPN);
  const float ry = __ldg(table + (size_t)row * LP_TABLE_STRIDE + LP_T_REPY);
  float xv[N], yv[N], nv[N];
  bool known[N];
  load_f<N>(x, i, xv);
  load_f_ro<N>(y, i, yv);
  load_f_ro<N>(noise, i, nv);
  load_m<N>(mask, mi, known);
#pragma unroll
  for (int j = 0; j < N; ++j) xv[j] = known[j] ? fmaf(rn, nv[j], ry * yv[j]) : xv[j];
  store_f<N>(x_model, i, xv);
  if (x_copy) store_f<N>(x_copy, i, xv);
}

template <int N>
__global__ void __launch_bounds__(kBlock) epilogue_kernel(const float* __restrict__ model_out,
                                                          const float* __restrict__ y,
                                                          const uint8_t* __restrict__ mask, float* out,
                                                          Geometry g) {
  pdl_prologue();
  const uint32_t i = (blockIdx.x * kBlock + threadIdx.x) * N;
  if (i >= g.total) return;
  uint32_t row, mi;
  locate(g, i, row, mi);
  float ov[N], yv[N];
  bool known[N];
  load_f_ro<N>(model_out, i, ov);
  load_f_ro<N>(y, i, yv);
  load_m<N>(mask, mi, known);
#pragma unroll
  for (int j = 0; j < N; ++j) ov[j] = known[j] ? yv[j] : ov[j];
  store_f<N>(out, i, ov);
}

template <int N>
__global__ void __launch_bounds__(kBlock) epilogue_euler_kernel(const float* __restrict__ model_out,
                                                                const float* __restrict__ y,
                                                                const uint8_t* __restrict__ mask, float* x,
                                                                float* out, float coef, Geometry g) {
  pdl_prologue();
  const uint32_t i = (blockIdx.x * kBlock + threadIdx.x) * N;
  if (i >= g.total) return;
  uint32_t row, mi;
  locate(g, i, row, mi);
  float ov[N], yv[N], xv[N];
  bool known[N];
  load_f_ro<N>(model_out, i, ov);
  load_f_ro<N>(y, i, yv);
  load_f<N>(x, i, xv);
  load_m<N>(mask, mi, known);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    ov[j] = known[j] ? yv[j] : ov[j];
    xv[j] = fmaf(xv[j] - ov[j], coef, xv[j]);
  }
  if (out) store_f<N>(out, i, ov);
  store_f<N>(x, i, xv);
}

template <int N>
__global__ void __launch_bounds__(kBlock) epilogue_cfg_kernel(const float* __restrict__ cond,
                                                              const float* __restrict__ uncond, float cfg,
                                                              const float* __restrict__ y,
                                                              const uint8_t* __restrict__ mask, float* x, float* out,
                                                              float coef, Geometry g) {
  pdl_prologue();
  const uint32_t i = (blockIdx.x * kBlock + threadIdx.x) * N;
  if (i >= g.total) return;
  uint32_t row, mi;
  locate(g, i, row, mi);
  float cv[N], uv[N], yv[N], xv[N];
  bool known[N];
  load_f_ro<N>(cond, i, cv);
  load_f_ro<N>(uncond, i, uv);
  load_f_ro<N>(y, i, yv);
  load_m<N>(mask, mi, known);
  if (x) load_f<N>(x, i, xv);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const float o = __fadd_rn(uv[j], __fmul_rn(__fsub_rn(cv[j], uv[j]), cfg));
    cv[j] = known[j] ? yv[j] : o;
    if (x) xv[j] = fmaf(xv[j] - cv[j], coef, xv[j]);
  }
  store_f<N>(out, i, cv);
  if (x) store_f<N>(x, i, xv);
}

template <int N>
__global__ void __launch_bounds__(kBlock) step_boundary_kernel(const float* __restrict__ model_out,
                                                               const float* __restrict__ y,
                                                               const float* __restrict__ noise,
                                                               const uint8_t* __restrict__ mask, float* x, float* out,
                                                               float coef, const float* __restrict__ next_table,
                                                               Geometry g) {
  pdl_prologue();
  const uint32_t i = (blockIdx.x * kBlock + threadIdx.x) * N;
  if (i >= g.total) return;
  uint32_t row, mi;
  locate(g, i, row, mi);
  bool known[N];
  load_m<N>(mask, mi, known);
  bool any_known = false, any_free = false;
#pragma unroll
  for (int j = 0; j < N; ++j) {
    any_known |= known[j];
    any_free |= !known[j];
  }
  float ov[N], yv[N], xv[N], nv[N];
#pragma unroll
  for (int j = 0; j < N; ++j) ov[j] = xv[j] = nv[j] = 0.f;
  load_f_ro<N>(y, i, yv);
  if (any_free) {  // free positions: denoised output and the running state
    load_f_ro<N>(model_out, i, ov);
    load_f<N>(x, i, xv);
  }
  if (any_known) load_f_ro<N>(noise, i, nv);  // known positions: re-noised copy of the clean latent
  const float rn = __ldg(next_table + (size_t)row * LP_TABLE_STRIDE + LP_T_REPN);
  const float ry = __ldg(next_table + (size_t)row * LP_TABLE_STRIDE + LP_T_REPY);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    const float o = known[j] ? yv[j] : ov[j];
    const float stepped = fmaf(xv[j] - o, coef, xv[j]);
    ov[j] = o;
    xv[j] = known[j] ? fmaf(rn, nv[j], ry * yv[j]) : stepped;
  }
  if (out) store_f<N>(out, i, ov);
  store_f<N>(x, i, xv);
}

__global__ void __launch_bounds__(kBlock) pack_mask_kernel(const float* __restrict__ m, uint8_t* out,
                                                           uint32_t n, int invert) {
  pdl_prologue();
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  const bool hi = __ldg(m + i) > 0.5f;
  out[i] = (hi != (invert != 0)) ? 1 : 0;
}

__global__ void __launch_bounds__(kBlock) fill_normal_philox_kernel(float* out, uint32_t n, uint64_t seed,
                                                                    uint64_t draw, const uint64_t* st) {
  pdl_prologue();
  const uint32_t i = blockIdx.x * kBlock + threadIdx.x;
  if (i >= n) return;
  if (st) {
    seed = st[0];
    draw += st[1];
  }
  out[i] = pick(philox_normal4(seed, draw, i >> 2), i & 3);
}

__global__ void __launch_bounds__(kBlock) fill_normal_torch_kernel(float* out, uint32_t n, uint64_t seed,
                                                                   uint64_t offset, const uint64_t* st,
                                                                   uint32_t T) {
  pdl_prologue();
  const uint32_t tid = blockIdx.x * kBlock + threadIdx.x;
  if (tid >= T) return;
  if (st) {
    seed = st[0];
    offset += st[1];
  }
  uint32_t call = 0;
  for (uint64_t base = tid; base < n; base += 4ull * T, ++call) {
    const float4 r = torch_normal4(seed, offset, tid, call);
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
      const uint64_t li = base + (uint64_t)ii * T;
      if (li < n) out[li] = pick(r, ii);
    }
  }
}

template <int N>
__global__ void __launch_bounds__(kBlock) synth_denoiser_kernel(const float* __restrict__ x, float* h0,
                                                                float* h1, uint32_t n, float a0, float b0,
                                                                float c0, float a1, float c1) {
  pdl_prologue();
  const uint32_t i = (blockIdx.x * kBlock + threadIdx.x) * N;
  if (i >= n) return;
  float xv[N], u[N], w[N];
  load_f_ro<N>(x, i, xv);
#pragma unroll
  for (int j = 0; j < N; ++j) {
    u[j] = fmaf(a0, xv[j], fmaf(b0, tanhf(xv[j]), c0));
    w[j] = fmaf(a1, xv[j], c1);
  }
  store_f<N>(h0, i, u);
  if (h1) store_f<N>(h1, i, w);
}

// ----------------------------------------------------------------------------
// host helpers
// ----------------------------------------------------------------------------
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }
inline bool aligned4(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 3u) == 0; }

inline int check_launch() {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    g_last_cuda_error = static_cast<int>(e);
    return LP_ERR_CUDA;
  }
  return LP_OK;
}

// Exact for every n < 2^31: with l = ceil(log2 d), mul = floor(2^(31+l)/d) + 1 < 2^32 and the error
// n * eps / 2^(31+l) stays below 1/d.  d == 1 is flagged with mul == 0.
inline FastDiv make_fastdiv(uint32_t d) {
  FastDiv f{d, 0u, 0u};
  if (d <= 1) return f;
  uint32_t l = 0;
  while ((1ull << l) < d) ++l;
  const unsigned shift = 31 + l;
  f.mul = static_cast<uint32_t>(((1ull << shift) / d) + 1);
  f.shift = l - 1;
  return f;
}

inline int make_geometry(const lp_dims* d, Geometry& g) {
  if (!d || d->n_rows < 0 || d->per_row <= 0 || d->spatial <= 0) return LP_ERR_INVALID;
  if (d->per_row % d->spatial != 0) return LP_ERR_INVALID;
  if (d->mask_row_stride < 0 || d->mask_channel_stride < 0) return LP_ERR_INVALID;
  const int64_t total = d->n_rows * d->per_row;
  const int64_t mask_extent = d->n_rows * d->mask_row_stride + d->per_row;  // loose upper bound
  if (total >= (int64_t(1) << 31) || mask_extent >= (int64_t(1) << 31)) return LP_ERR_UNSUPPORTED;
  g.total = static_cast<uint32_t>(total);
  g.per_row = make_fastdiv(static_cast<uint32_t>(d->per_row));
  g.spatial = make_fastdiv(static_cast<uint32_t>(d->spatial));
  g.mask_row_stride = static_cast<uint32_t>(d->mask_row_stride);
  g.mask_channel_stride = static_cast<uint32_t>(d->mask_channel_stride);
  if (d->row_split < 0 || d->row_split > d->per_row) return LP_ERR_INVALID;
  g.n_rows = static_cast<uint32_t>(d->n_rows);
  g.row_split = static_cast<uint32_t>(d->row_split);
  return LP_OK;
}

// 128-bit path needs every row / channel / mask offset to stay 4-aligned.
inline bool geometry_vec4(const Geometry& g, const uint8_t* mask) {
  return g.per_row.d % 4 == 0 && g.spatial.d % 4 == 0 && g.row_split % 4 == 0 && g.mask_row_stride % 4 == 0 &&
         g.mask_channel_stride % 4 == 0 && aligned4(mask);
}

inline unsigned blocks_for(uint32_t n_threads) { return (n_threads + kBlock - 1) / kBlock; }

int torch_grid(int64_t numel, int device, int64_t* grid, uint64_t* inc) {
  if (numel <= 0) return LP_ERR_INVALID;
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) return LP_ERR_CUDA;
  }
  int sms = 0, tpsm = 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess ||
      cudaDeviceGetAttribute(&tpsm, cudaDevAttrMaxThreadsPerMultiProcessor, device) != cudaSuccess) {
    g_last_cuda_error = static_cast<int>(cudaGetLastError());
    return LP_ERR_CUDA;
  }
  int64_t g = (numel + 255) / 256;
  const int64_t cap = int64_t(sms) * (tpsm / 256);
  if (g > cap) g = cap;
  if (grid) *grid = g;
  if (inc) *inc = (uint64_t)(((numel - 1) / (256 * g * 4) + 1) * 4);
  return LP_OK;
}

inline int tma_mode() { return g_opt_tma; }  // 1 (default) = when eligible, 0 = never (see lp_set_option)

// The TMA-staged variant needs 16-byte aligned slices: spatial a multiple of 16 (mask slices) and no
// side outputs; only worth it when there are enough tiles to keep a persistent grid busy.
template <bool kFirst, bool kNext, bool kMerge, int kTile, int kStages, int kMinBlocks>
int launch_substep_tma_cfg(const SubstepArgs& a, cudaStream_t s) {
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  }
  const size_t smem = sizeof(TmaStage<kTile>) * kStages;
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(substep_tma_kernel<kFirst, kNext, kMerge, kTile, kStages, kMinBlocks>,
                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    configured = true;
  }
  TileGeom tg;
  tg.channels = a.g.per_row.d / a.g.spatial.d;
  tg.tiles_per_channel = (a.g.spatial.d + kTile - 1) / kTile;
  tg.n_tiles = (a.g.total / a.g.spatial.d) * tg.tiles_per_channel;
  unsigned grid = (unsigned)sms * (unsigned)kMinBlocks;
  if (grid > tg.n_tiles) grid = tg.n_tiles;
  launch_kernel_smem(substep_tma_kernel<kFirst, kNext, kMerge, kTile, kStages, kMinBlocks>, dim3(grid), smem, s, a, tg);
  return check_launch();
}

// tile / ring geometry: "tma" option value 1 (default) = 2048-element tiles, 2 stages, 2 CTAs per SM;
// 2, 3, 4 are the alternatives measured in profiles/README.md
template <bool kFirst, bool kNext, bool kMerge>
int launch_substep_tma(const SubstepArgs& a, cudaStream_t s) {
  if (kNext && !kFirst && kMerge) {  // the steady kernel: alternative geometries stay selectable for measurement
    switch (tma_mode()) {
      case 2: return launch_substep_tma_cfg<kFirst, kNext, kMerge, 1024, 4, 2>(a, s);
      case 3: return launch_substep_tma_cfg<kFirst, kNext, kMerge, 4096, 2, 1>(a, s);
      case 4: return launch_substep_tma_cfg<kFirst, kNext, kMerge, 2048, 4, 1>(a, s);
      case 5: return launch_substep_tma_cfg<kFirst, kNext, kMerge, 1024, 3, 3>(a, s);
      default: break;
    }
  }
  return launch_substep_tma_cfg<kFirst, kNext, kMerge, 2048, 2, 2>(a, s);
}

inline bool tma_eligible(const SubstepArgs& a) {
  return tma_mode() != 0 && a.g.spatial.d % 16 == 0 && a.g.mask_row_stride % 16 == 0 && a.g.mask_channel_stride % 16 == 0 &&
         (reinterpret_cast<uintptr_t>(a.mask) & 15u) == 0 && a.g.row_split == 0 && !a.x_copy && !a.x0e &&
         a.g.total >= (1u << 20);
}

template <int N, int kRng>
int launch_substep_vec(const SubstepArgs& a, bool first, bool next, bool merge, cudaStream_t s) {
  const unsigned grid = blocks_for((a.g.total + N - 1) / N);
  if (N == 4 && kRng == LP_RNG_PHILOX && tma_eligible(a)) {
    if (first && next && merge) return launch_substep_tma<true, true, true>(a, s);
    if (first && next) return launch_substep_tma<true, true, false>(a, s);
    if (next && merge) return launch_substep_tma<false, true, true>(a, s);
    if (next) return launch_substep_tma<false, true, false>(a, s);
    if (first) return launch_substep_tma<true, false, false>(a, s);
    return launch_substep_tma<false, false, false>(a, s);
  }
  if (merge && kRng == LP_RNG_PHILOX) {
    if (first) launch_kernel(substep_kernel<N, LP_RNG_PHILOX, true, true, true>, dim3(grid), s, a);
    else launch_kernel(substep_kernel<N, LP_RNG_PHILOX, false, true, true>, dim3(grid), s, a);
    return check_launch();
  }
  if (first && next) launch_kernel(substep_kernel<N, kRng, true, true>, dim3(grid), s, a);
  else if (first) launch_kernel(substep_kernel<N, kRng, true, false>, dim3(grid), s, a);
  else if (next) launch_kernel(substep_kernel<N, kRng, false, true>, dim3(grid), s, a);
  else launch_kernel(substep_kernel<N, kRng, false, false>, dim3(grid), s, a);
  return check_launch();
}

}  // namespace lp

// ============================================================================
// C ABI
// ============================================================================
using namespace lp;

extern "C" int lp_abi_version(void) { return LP_ABI_VERSION; }

extern "C" const char* lp_status_string(int status) {
  switch (status) {
    case LP_OK: return "ok";
    case LP_ERR_INVALID: return "invalid argument";
    case LP_ERR_ALIGNMENT: return "misaligned pointer";
    case LP_ERR_CUDA: return "CUDA launch failed (see lp_last_cuda_error)";
    case LP_ERR_UNSUPPORTED: return "unsupported size (>= 2^31 elements)";
    default: return "unknown status";
  }
}

extern "C" int lp_last_cuda_error(void) { return g_last_cuda_error; }

extern "C" int lp_set_option(const char* name, int value) {
  if (!name) return LP_ERR_INVALID;
  const auto eq = [&](const char* k) { int i = 0; while (k[i] && name[i] == k[i]) ++i; return !k[i] && !name[i]; };
  if (eq("pdl")) { g_opt_pdl = value; return LP_OK; }
  if (eq("tma")) { g_opt_tma = value; return LP_OK; }
  return LP_ERR_INVALID;
}

extern "C" int64_t lp_selftest_index_math(int64_t samples) {
  int64_t bad = 0;
  auto check = [&](uint32_t n, uint32_t d) {
    const FastDiv f = make_fastdiv(d);
    const uint32_t q = f.mul ? (uint32_t)((((uint64_t)n * f.mul) >> 32) >> f.shift) : n;  // == FastDiv::div
    if (q != n / d) ++bad;
  };
  const uint32_t ds[] = {1u, 2u, 3u, 4u, 5u, 7u, 35u, 105u, 1024u, 3600u, 16384u, 65536u, 75600u, 1209600u,
                         (1u << 30), (1u << 31) - 1u, 0x7fffffffu, 1000003u};
  const uint32_t ns[] = {0u, 1u, 2u, 3u, 1023u, 65535u, 65536u, (1u << 31) - 1u, (1u << 31) - 2u, 0x40000000u};
  for (uint32_t d : ds)
    for (uint32_t n : ns) {
      check(n, d);
      if (d > 1 && n >= d) { check(n - n % d, d); check(n - n % d - 1, d); }
    }
  uint64_t st = 0x9E3779B97F4A7C15ull;
  for (int64_t k = 0; k < samples; ++k) {
    st = st * 6364136223846793005ull + 1442695040888963407ull;
    const uint32_t n = (uint32_t)(st >> 33);                       // < 2^31
    st = st * 6364136223846793005ull + 1442695040888963407ull;
    uint32_t d = (uint32_t)(st >> 33) >> ((st >> 20) & 31);        // all magnitudes
    if (d == 0) d = 1;
    check(n, d);
  }
  return bad;
}

extern "C" int lp_torch_randn_geometry(int64_t numel, int device, int64_t* grid_out, uint64_t* increment_out) {
  return torch_grid(numel, device, grid_out, increment_out);
}

extern "C" int lp_pack_mask_f32(const float* mask_f32, uint8_t* mask_u8, int64_t n, int invert,
                                lp_stream_t stream) {
  if (!mask_f32 || !mask_u8 || n < 0) return LP_ERR_INVALID;
  if (n >= (int64_t(1) << 31)) return LP_ERR_UNSUPPORTED;
  if (n == 0) return LP_OK;
  launch_kernel(pack_mask_kernel, dim3(blocks_for((uint32_t)n)), (cudaStream_t)stream, mask_f32, mask_u8,
                                                                                 (uint32_t)n, invert);
  return check_launch();
}

extern "C" int lp_prologue_f32(const float* x, const float* y, const float* noise, const uint8_t* mask,
                               float* x_model, float* x_copy, const float* table, const lp_dims* dims,
                               lp_stream_t stream) {
  if (!x || !y || !noise || !mask || !x_model || !table) return LP_ERR_INVALID;
  Geometry g;
  if (int rc = make_geometry(dims, g)) return rc;
  if (g.total == 0) return LP_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const bool v4 = geometry_vec4(g, mask) && aligned16(x) && aligned16(y) && aligned16(noise) &&
                  aligned16(x_model) && (!x_copy || aligned16(x_copy));
  if (v4) launch_kernel(prologue_kernel<4>, dim3(blocks_for(g.total / 4)), s, x, y, noise, mask, x_model, x_copy, table, g);
  else launch_kernel(prologue_kernel<1>, dim3(blocks_for(g.total)), s, x, y, noise, mask, x_model, x_copy, table, g);
  return check_launch();
}

static int substep_impl(float* x_model, const float* x0, const float* x0_big, const float* y,
                        const uint8_t* mask, float* c_state, float* x_copy, float* x0e_out, const float* table,
                        const lp_dims* dims, const lp_rng* rng, int flags, lp_stream_t stream, int use_cfg,
                        float cfg, float cfg_big) {
  if (!x_model || !x0 || !y || !mask || !table || !rng) return LP_ERR_INVALID;
  if (flags & ~(LP_SUBSTEP_FIRST | LP_SUBSTEP_FUSE_NEXT | LP_SUBSTEP_STORE_C | LP_SUBSTEP_MERGE_NOISE))
    return LP_ERR_INVALID;
  const bool merge = (flags & LP_SUBSTEP_MERGE_NOISE) != 0;
  if (merge && (!(flags & LP_SUBSTEP_FUSE_NEXT) || rng->mode != LP_RNG_PHILOX)) return LP_ERR_INVALID;
  const int first = (flags & LP_SUBSTEP_FIRST) != 0;
  const int has_next = (flags & LP_SUBSTEP_FUSE_NEXT) != 0;
  if (!x0_big) x0_big = x0;
  if ((!first || has_next || (flags & LP_SUBSTEP_STORE_C)) && !c_state) return LP_ERR_INVALID;
  SubstepArgs a;
  if (int rc = make_geometry(dims, a.g)) return rc;
  if (a.g.total == 0) return LP_OK;
  a.x = x_model; a.x0 = x0; a.x0b = x0_big; a.y = y; a.mask = mask; a.c = c_state;
  a.x_copy = x_copy; a.x0e = x0e_out; a.table = table;
  a.tape0 = rng->tape0; a.tape1 = rng->tape1; a.rng_state = rng->state;
  a.seed = rng->seed; a.draw0 = rng->draw0; a.draw1 = rng->draw1; a.torch_T = 0;
  a.store_c = (flags & LP_SUBSTEP_STORE_C) != 0;
  a.use_cfg = use_cfg; a.cfg = cfg; a.cfg_big = cfg_big;
  cudaStream_t s = (cudaStream_t)stream;
  const bool f = first != 0, n = has_next != 0;

  if (rng->mode == LP_RNG_TORCH) {
    int64_t grid = 0;
    if (int rc = torch_grid(a.g.total, -1, &grid, nullptr)) return rc;
    a.torch_T = (uint32_t)(grid * 256);
    const unsigned gb = (unsigned)grid;
    const bool tv4 = geometry_vec4(a.g, mask) && aligned16(x_model) && aligned16(x0) && aligned16(x0_big) &&
                     aligned16(y) && (!c_state || aligned16(c_state)) && (!x_copy || aligned16(x_copy)) &&
                     (!x0e_out || aligned16(x0e_out));
    const uint64_t calls = ((uint64_t)a.g.total + 4ull * a.torch_T - 1) / (4ull * a.torch_T);
    if (tv4 && calls <= 65535) {
      const dim3 g2(gb, (unsigned)calls);
      if (f && n) launch_kernel(substep_torchvec_kernel<true, true>, dim3(g2), s, a);
      else if (f) launch_kernel(substep_torchvec_kernel<true, false>, dim3(g2), s, a);
      else if (n) launch_kernel(substep_torchvec_kernel<false, true>, dim3(g2), s, a);
      else launch_kernel(substep_torchvec_kernel<false, false>, dim3(g2), s, a);
      return check_launch();
    }
    if (f && n) launch_kernel(substep_torch_kernel<true, true>, dim3(gb), s, a);
    else if (f) launch_kernel(substep_torch_kernel<true, false>, dim3(gb), s, a);
    else if (n) launch_kernel(substep_torch_kernel<false, true>, dim3(gb), s, a);
    else launch_kernel(substep_torch_kernel<false, false>, dim3(gb), s, a);
    return check_launch();
  }

  bool v4 = geometry_vec4(a.g, mask) && aligned16(x_model) && aligned16(x0) && aligned16(x0_big) &&
            aligned16(y) && (!c_state || aligned16(c_state)) && (!x_copy || aligned16(x_copy)) &&
            (!x0e_out || aligned16(x0e_out));
  if (rng->mode == LP_RNG_TAPE) {
    if (!rng->tape0 || (n && !rng->tape1)) return LP_ERR_INVALID;
    v4 = v4 && aligned16(rng->tape0) && (!n || aligned16(rng->tape1));
    return v4 ? launch_substep_vec<4, LP_RNG_TAPE>(a, f, n, false, s) : launch_substep_vec<1, LP_RNG_TAPE>(a, f, n, false, s);
  }
  if (rng->mode == LP_RNG_PHILOX) {
    return v4 ? launch_substep_vec<4, LP_RNG_PHILOX>(a, f, n, merge, s)
              : launch_substep_vec<1, LP_RNG_PHILOX>(a, f, n, merge, s);
  }
  return LP_ERR_INVALID;
}

extern "C" int lp_substep_f32(float* x_model, const float* x0, const float* x0_big, const float* y,
                              const uint8_t* mask, float* c_state, float* x_copy, float* x0e_out,
                              const float* table, const lp_dims* dims, const lp_rng* rng, int flags,
                              lp_stream_t stream) {
  return substep_impl(x_model, x0, x0_big, y, mask, c_state, x_copy, x0e_out, table, dims, rng, flags, stream, 0,
                      0.f, 0.f);
}

extern "C" int lp_substep_cfg_f32(float* x_model, const float* cond, const float* uncond, float cfg, float cfg_big,
                                  const float* y, const uint8_t* mask, float* c_state, float* x_copy,
                                  float* x0e_out, const float* table, const lp_dims* dims, const lp_rng* rng,
                                  int flags, lp_stream_t stream) {
  if (!uncond || uncond == cond) return LP_ERR_INVALID;
  return substep_impl(x_model, cond, uncond, y, mask, c_state, x_copy, x0e_out, table, dims, rng, flags, stream, 1,
                      cfg, cfg_big);
}

extern "C" int lp_step_boundary_f32(const float* model_out, const float* y, const float* noise, const uint8_t* mask,
                                    float* x_inout, float* out, float euler_coef, const float* next_table,
                                    const lp_dims* dims, lp_stream_t stream) {
  if (!model_out || !y || !noise || !mask || !x_inout || !next_table) return LP_ERR_INVALID;
  Geometry g;
  if (int rc = make_geometry(dims, g)) return rc;
  if (g.total == 0) return LP_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const bool v4 = geometry_vec4(g, mask) && aligned16(model_out) && aligned16(y) && aligned16(noise) &&
                  aligned16(x_inout) && (!out || aligned16(out));
  if (v4) launch_kernel(step_boundary_kernel<4>, dim3(blocks_for(g.total / 4)), s, model_out, y, noise, mask, x_inout, out, euler_coef, next_table, g);
  else launch_kernel(step_boundary_kernel<1>, dim3(blocks_for(g.total)), s, model_out, y, noise, mask, x_inout, out, euler_coef, next_table, g);
  return check_launch();
}

extern "C" int lp_epilogue_cfg_f32(const float* cond, const float* uncond, float cfg, const float* y,
                                   const uint8_t* mask, float* x_inout, float* out, float euler_coef,
                                   const lp_dims* dims, lp_stream_t stream) {
  if (!cond || !uncond || !y || !mask || !out) return LP_ERR_INVALID;
  Geometry g;
  if (int rc = make_geometry(dims, g)) return rc;
  if (g.total == 0) return LP_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const bool v4 = geometry_vec4(g, mask) && aligned16(cond) && aligned16(uncond) && aligned16(y) && aligned16(out) &&
                  (!x_inout || aligned16(x_inout));
  if (v4) launch_kernel(epilogue_cfg_kernel<4>, dim3(blocks_for(g.total / 4)), s, cond, uncond, cfg, y, mask, x_inout, out, euler_coef, g);
  else launch_kernel(epilogue_cfg_kernel<1>, dim3(blocks_for(g.total)), s, cond, uncond, cfg, y, mask, x_inout, out, euler_coef, g);
  return check_launch();
}

extern "C" int lp_advance_f32(float* x_model, const float* c_state, const uint8_t* mask, const float* table,
                              const lp_dims* dims, const lp_rng* rng, int half, lp_stream_t stream) {
  if (!x_model || !c_state || !mask || !table || !rng) return LP_ERR_INVALID;
  AdvanceArgs a;
  if (int rc = make_geometry(dims, a.g)) return rc;
  if (a.g.total == 0) return LP_OK;
  a.x = x_model; a.c = c_state; a.mask = mask; a.table = table; a.tape0 = rng->tape0;
  a.rng_state = rng->state; a.seed = rng->seed; a.draw0 = rng->draw0; a.torch_T = 0; a.half = half != 0;
  cudaStream_t s = (cudaStream_t)stream;
  if (rng->mode == LP_RNG_TORCH) {
    int64_t grid = 0;
    if (int rc = torch_grid(a.g.total, -1, &grid, nullptr)) return rc;
    a.torch_T = (uint32_t)(grid * 256);
    launch_kernel(advance_torch_kernel, dim3((unsigned)grid), s, a);
  } else if (rng->mode == LP_RNG_TAPE) {
    if (!rng->tape0) return LP_ERR_INVALID;
    launch_kernel(advance_kernel<LP_RNG_TAPE>, dim3(blocks_for(a.g.total)), s, a);
  } else if (rng->mode == LP_RNG_PHILOX) {
    launch_kernel(advance_kernel<LP_RNG_PHILOX>, dim3(blocks_for(a.g.total)), s, a);
  } else {
    return LP_ERR_INVALID;
  }
  return check_launch();
}

extern "C" int lp_epilogue_f32(const float* model_out, const float* y, const uint8_t* mask, float* out,
                               const lp_dims* dims, lp_stream_t stream) {
  if (!model_out || !y || !mask || !out) return LP_ERR_INVALID;
  Geometry g;
  if (int rc = make_geometry(dims, g)) return rc;
  if (g.total == 0) return LP_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const bool v4 = geometry_vec4(g, mask) && aligned16(model_out) && aligned16(y) && aligned16(out);
  if (v4) launch_kernel(epilogue_kernel<4>, dim3(blocks_for(g.total / 4)), s, model_out, y, mask, out, g);
  else launch_kernel(epilogue_kernel<1>, dim3(blocks_for(g.total)), s, model_out, y, mask, out, g);
  return check_launch();
}

extern "C" int lp_epilogue_euler_f32(const float* model_out, const float* y, const uint8_t* mask, float* x_inout,
                                     float* out, float euler_coef, const lp_dims* dims, lp_stream_t stream) {
  if (!model_out || !y || !mask || !x_inout) return LP_ERR_INVALID;
  Geometry g;
  if (int rc = make_geometry(dims, g)) return rc;
  if (g.total == 0) return LP_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const bool v4 = geometry_vec4(g, mask) && aligned16(model_out) && aligned16(y) && (!out || aligned16(out)) &&
                  aligned16(x_inout);
  if (v4) launch_kernel(epilogue_euler_kernel<4>, dim3(blocks_for(g.total / 4)), s, model_out, y, mask, x_inout, out, euler_coef, g);
  else launch_kernel(epilogue_euler_kernel<1>, dim3(blocks_for(g.total)), s, model_out, y, mask, x_inout, out, euler_coef, g);
  return check_launch();
}

extern "C" int lp_stop_stats_f32(const float* a, const float* b, const uint8_t* mask, const uint8_t* ring,
                                 const float* table, const lp_dims* dims, double* sums, lp_stream_t stream) {
  if (!a || !b || !mask || !sums) return LP_ERR_INVALID;
  StatsArgs s;
  if (int rc = make_geometry(dims, s.g)) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  if (cudaMemsetAsync(sums, 0, 2 * sizeof(double), st) != cudaSuccess) return check_launch() ? LP_ERR_CUDA : LP_ERR_CUDA;
  if (s.g.total == 0) return LP_OK;
  s.a = a; s.b = b; s.mask = mask; s.ring = ring; s.table = table; s.sums = sums;
  int sms = 148;
  int device = 0;
  if (cudaGetDevice(&device) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  const bool v4 = geometry_vec4(s.g, mask) && aligned16(a) && aligned16(b) && (!ring || aligned4(ring));
  const uint32_t groups = v4 ? s.g.total / 4 : s.g.total;
  unsigned grid = blocks_for(groups);
  const unsigned cap = (unsigned)sms * 8;  // persistent-sized: 8 CTAs of 256 threads per SM
  if (grid > cap) grid = cap;
  if (v4) launch_kernel(stop_stats_kernel<4>, dim3(grid), st, s);
  else launch_kernel(stop_stats_kernel<1>, dim3(grid), st, s);
  return check_launch();
}

extern "C" int lp_fill_normal_f32(float* out, int64_t n, const lp_rng* rng, lp_stream_t stream) {
  if (!out || !rng || n < 0) return LP_ERR_INVALID;
  if (n >= (int64_t(1) << 31)) return LP_ERR_UNSUPPORTED;
  if (n == 0) return LP_OK;
  cudaStream_t s = (cudaStream_t)stream;
  if (rng->mode == LP_RNG_PHILOX) {
    launch_kernel(fill_normal_philox_kernel, dim3(blocks_for((uint32_t)n)), s, out, (uint32_t)n, rng->seed,
                                                                        rng->draw0, rng->state);
    return check_launch();
  }
  if (rng->mode == LP_RNG_TORCH) {
    int64_t grid = 0;
    if (int rc = torch_grid(n, -1, &grid, nullptr)) return rc;
    launch_kernel(fill_normal_torch_kernel, dim3((unsigned)grid), s, out, (uint32_t)n, rng->seed, rng->draw0,
                                                              rng->state, (uint32_t)(grid * 256));
    return check_launch();
  }
  return LP_ERR_INVALID;
}

extern "C" int lp_synth_denoiser_f32(const float* x, float* h0, float* h1, int64_t n, const float* coef,
                                     lp_stream_t stream) {
  if (!x || !h0 || !coef || n < 0) return LP_ERR_INVALID;
  if (n >= (int64_t(1) << 31)) return LP_ERR_UNSUPPORTED;
  if (n == 0) return LP_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const bool v4 = n % 4 == 0 && aligned16(x) && aligned16(h0) && (!h1 || aligned16(h1));
  if (v4)
    launch_kernel(synth_denoiser_kernel<4>, dim3(blocks_for((uint32_t)n / 4)), s, x, h0, h1, (uint32_t)n, coef[0],
                                                                          coef[1], coef[2], coef[3], coef[4]);
  else
    launch_kernel(synth_denoiser_kernel<1>, dim3(blocks_for((uint32_t)n)), s, x, h0, h1, (uint32_t)n, coef[0], coef[1],
                                                                      coef[2], coef[3], coef[4]);
  return check_launch();
}

extern "C" int lp_l2_persist_capacity(int device, size_t* max_persisting_bytes, size_t* max_window_bytes) {
  if (device < 0 && cudaGetDevice(&device) != cudaSuccess) return LP_ERR_CUDA;
  int a = 0, b = 0;
  if (cudaDeviceGetAttribute(&a, cudaDevAttrMaxPersistingL2CacheSize, device) != cudaSuccess ||
      cudaDeviceGetAttribute(&b, cudaDevAttrMaxAccessPolicyWindowSize, device) != cudaSuccess) {
    g_last_cuda_error = static_cast<int>(cudaGetLastError());
    return LP_ERR_CUDA;
  }
  if (max_persisting_bytes) *max_persisting_bytes = (size_t)a;
  if (max_window_bytes) *max_window_bytes = (size_t)b;
  return LP_OK;
}

extern "C" int lp_l2_persist_set(const void* ptr, size_t bytes, lp_stream_t stream) {
  if (!ptr || bytes == 0) return LP_ERR_INVALID;
  size_t cap = 0, win = 0;
  if (int rc = lp_l2_persist_capacity(-1, &cap, &win)) return rc;
  if (cap == 0 || win == 0) return LP_ERR_UNSUPPORTED;
  const size_t want = bytes < cap ? bytes : cap;
  if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) != cudaSuccess) return check_launch() ? LP_ERR_CUDA : LP_ERR_CUDA;
  cudaStreamAttrValue v;
  v.accessPolicyWindow.base_ptr = const_cast<void*>(ptr);
  v.accessPolicyWindow.num_bytes = bytes < win ? bytes : win;
  v.accessPolicyWindow.hitRatio = (float)((double)want / (double)(bytes < win ? bytes : win));
  if (v.accessPolicyWindow.hitRatio > 1.f) v.accessPolicyWindow.hitRatio = 1.f;
  v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
  v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
  if (cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess)
    return check_launch() ? LP_ERR_CUDA : LP_ERR_CUDA;
  return LP_OK;
}

extern "C" int lp_l2_persist_clear(lp_stream_t stream) {
  cudaStreamAttrValue v;
  v.accessPolicyWindow.base_ptr = nullptr;
  v.accessPolicyWindow.num_bytes = 0;
  v.accessPolicyWindow.hitRatio = 0.f;
  v.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
  v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
  cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &v);
  if (cudaCtxResetPersistingL2Cache() != cudaSuccess) return check_launch() ? LP_ERR_CUDA : LP_ERR_CUDA;
  cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, 0);  // give the set-aside back to normal traffic
  return LP_OK;
}

extern "C" int lp_l2_flush(void* scratch, size_t bytes, lp_stream_t stream) {
  if (!scratch) return LP_ERR_INVALID;
  if (cudaMemsetAsync(scratch, 0, bytes, (cudaStream_t)stream) != cudaSuccess) return check_launch() ? LP_ERR_CUDA : LP_ERR_CUDA;
  return LP_OK;
}
