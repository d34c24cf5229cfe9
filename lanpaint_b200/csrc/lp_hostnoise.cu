// lanpaint_b200: the noise image of a sample call, drawn on the device with the bits of ComfyUI's CPU draw.
//
// ComfyUI's `comfy.sample.prepare_noise` (called by nodes.common_ksampler, i.e. by every LanPaint KSampler node:
// src/LanPaint/nodes.py:513,589) is `torch.manual_seed(seed); torch.randn(latent.size(), generator=..., device="cpu")`.
// At a batch of 128 SDXL latents that single-threaded CPU draw is 29 ms of a 38 ms call.  Its stream is fully
// specified by torch's CPU sources, so it can be produced here instead:
//
//   * the generator: at::mt19937 seeded with init_genrand(seed & 0xffffffff); one 32-bit output per float,
//     u = (y & 0xffffff) * 2^-24                              (ATen/core/MT19937RNGEngine.h, TransformationHelper.h);
//   * the transform: normal_fill (ATen/native/cpu/DistributionTemplates.h) -- every 16 consecutive uniforms become 16
//     normals by Box-Muller over the pairs (j, j+8): r = sqrt(-2 log(1-u_j)), theta = 2pi u_{j+8}, out_j = r cos theta,
//     out_{j+8} = r sin theta; a size that is not a multiple of 16 redraws its last 16 values from 16 fresh uniforms;
//   * the arithmetic: on every x86 build with AVX2 (the AVX512 dispatch falls back to the AVX2 kernel) log / sincos are
//     avx_mathfun.h's single-precision cephes routines, compiled with FMA contraction.  They are restated below with
//     every fused / unfused operation written out (`__fmaf_rn` / `__fmul_rn` / `__fadd_rn`), which reproduces the CPU
//     result bit for bit (checked against torch.randn on the GPU box by tests/test_gpu_hostnoise.py; the Python side
//     self-checks once per process and leaves ComfyUI's own function in place if the host's torch draws other bits).
//
// MT19937 is one sequential recurrence, x[n] = x[n-227] ^ twist(x[n-624], x[n-623]).  Column c of a "row" of 227
// consecutive words depends on the same column of the previous row (same thread, a register) and on two words 2.7
// rows back (other threads): ONE CTA walks the stream with 227 threads, the last 2048 words in a shared-memory ring,
// one barrier per TWO rows, and stores the raw words; tempering, the conversion to a uniform and the Box-Muller
// transform are a second, fully parallel launch.
#include "lp_common.cuh"

namespace lp {

constexpr int kMtN = 624, kMtM = 397, kMtLag = kMtN - kMtM;  // 227
constexpr int kMtRing = 2048;                                  // power of two; see the hazard argument below
constexpr int kMtMask = kMtRing - 1;
constexpr int kMtThreads = 256;                                // 227 of them walk the recurrence
constexpr int kMtSpan = 2 * kMtLag;                            // words per barrier interval (two rows)

__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v) {
  const uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
  return (y >> 1) ^ ((0u - (v & 1u)) & 0x9908b0dfu);
}

// tempering + at::uniform_real_distribution<float>: 24 bits of the output times 2^-24 (exact)
__device__ __forceinline__ float mt_uniform(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return __uint2float_rn(y & 0x00ffffffu) * 5.9604644775390625e-08f;
}

// raw[i] = the i-th generator word BEFORE tempering, for i < n_out (the parallel transform kernel tempers: nothing
// that can be done elsewhere stays on the sequential path); words are generated through n_words (a multiple of 624,
// >= n_out) so that state_out receives the engine's complete state array after its last twist.
//
// With X = i + 624 the index of word i in the sequence that starts with the seeded state x[0..623], at::mt19937 is
// x[X] = x[X-227] ^ twist(x[X-624], x[X-623]).  Thread c owns column c of every "row" of 227 words: x[X-227] is its own
// previous result (a register), x[X-624] / x[X-623] were written two to three rows earlier by other threads.  So two rows
// are produced between barriers.  Ring of 2048 words: within one interval the threads write x-indices
// [624 + 454 t, 624 + 454 t + 453] and read [454 t, 454 t + 454]: no two of these are 2048 apart, and a slot is
// overwritten 4.5 intervals after it was written.
__global__ void __launch_bounds__(kMtThreads, 1) mt19937_raw_kernel(uint32_t* __restrict__ raw, int64_t n_out,
                                                                      int64_t n_words, uint32_t seed,
                                                                      uint32_t* __restrict__ state_out) {
  __shared__ uint32_t ring[kMtRing];
  pdl_prologue();
  const int c = threadIdx.x;
  if (c == 0) {  // init_genrand: the engine's state before its first twist is x[0..623]
    uint32_t s = seed;
    ring[0] = s;
    for (int j = 1; j < kMtN; ++j) {
      s = 1812433253u * (s ^ (s >> 30)) + (uint32_t)j;
      ring[j] = s;
    }
  }
  __syncthreads();
  const int64_t intervals = (n_words + kMtSpan - 1) / kMtSpan;
  const int64_t state_lo = n_words - kMtN;
  // intervals that lie entirely inside raw[0, n_out) and before the state block: no bounds to test in the hot loop
  int64_t fast = n_out / kMtSpan;
  if (state_out && state_lo / kMtSpan < fast) fast = state_lo / kMtSpan;
  // threads 227..255 shadow column 226 (same loads, same values, no stores): the hot loop has no divergent branch
  const int col = c < kMtLag ? c : kMtLag - 1;
  const bool owner = c < kMtLag;
  uint32_t prev = ring[kMtM + col];                        // x[X-227] of row 0
  uint32_t p = (uint32_t)col;                              // ring slot of x[X-624] of this interval's first row
  const uint32_t base = smem_u32(ring);                    // shared-window address, computed once
  uint32_t* dst = raw + col;
  int64_t t = 0;
  for (int64_t left = fast; left > 0; left -= (left > 0x40000000 ? 0x40000000 : left)) {
    const int chunk = (int)(left > 0x40000000 ? 0x40000000 : left);   // 32-bit trip count for the hot loop
#pragma unroll 1
    for (int k = 0; k < chunk; ++k) {
      uint32_t a0, b0, a1, b1;
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a0) : "r"(base + 4u * p));
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(b0) : "r"(base + 4u * ((p + 1) & kMtMask)));
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(a1) : "r"(base + 4u * ((p + kMtLag) & kMtMask)));
      asm volatile("ld.shared.u32 %0, [%1];" : "=r"(b1) : "r"(base + 4u * ((p + kMtLag + 1) & kMtMask)));
      const uint32_t x0 = prev ^ mt_twist(a0, b0);
      const uint32_t x1 = x0 ^ mt_twist(a1, b1);
      if (owner) {   // predicated stores, no divergence: the shadow lanes only keep the warp converged
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + 4u * ((p + kMtN) & kMtMask)), "r"(x0) : "memory");
        asm volatile("st.shared.u32 [%0], %1;" ::"r"(base + 4u * ((p + kMtN + kMtLag) & kMtMask)), "r"(x1) : "memory");
        dst[0] = x0;
        dst[kMtLag] = x1;
      }
      prev = x1;
      p = (p + kMtSpan) & kMtMask;
      dst += kMtSpan;
      __syncthreads();
    }
    t += chunk;
  }
  const bool active = owner;
#pragma unroll 1
  for (; t < intervals; ++t) {   // the last few intervals: the end of raw[], the engine's final state block
    if (active) {
      const uint32_t a0 = ring[p], b0 = ring[(p + 1) & kMtMask];
      const uint32_t a1 = ring[(p + kMtLag) & kMtMask], b1 = ring[(p + kMtLag + 1) & kMtMask];
      const uint32_t x0 = prev ^ mt_twist(a0, b0);
      const uint32_t x1 = x0 ^ mt_twist(a1, b1);
      ring[(p + kMtN) & kMtMask] = x0;
      ring[(p + kMtN + kMtLag) & kMtMask] = x1;
      const int64_t i0 = t * kMtSpan + c, i1 = i0 + kMtLag;
      if (i0 < n_out) raw[i0] = x0;
      if (i1 < n_out) raw[i1] = x1;
      if (state_out) {
        if (i0 >= state_lo && i0 < n_words) state_out[i0 - state_lo] = x0;
        if (i1 >= state_lo && i1 < n_words) state_out[i1 - state_lo] = x1;
      }
      prev = x1;
      p = (p + kMtSpan) & kMtMask;
    }
    __syncthreads();
  }
}

// ---- avx_mathfun.h's log256_ps / sincos256_ps as torch's AVX2 build executes them, one lane ---------------------
__device__ __forceinline__ float cephes_logf_avx(float x) {
  x = fmaxf(x, __uint_as_float(0x00800000u));
  const uint32_t bits = __float_as_uint(x);
  float e = __fadd_rn(__int2float_rn((int)(bits >> 23) - 0x7f), 1.0f);
  x = __uint_as_float((bits & ~0x7f800000u) | 0x3f000000u);                 // mantissa in [0.5, 1)
  const bool lt = x < 0.707106781186547524f;
  const float tmp = lt ? x : 0.0f;
  x = __fsub_rn(x, 1.0f);
  e = __fsub_rn(e, lt ? 1.0f : 0.0f);
  x = __fadd_rn(x, tmp);
  const float z = __fmul_rn(x, x);
  float y = 7.0376836292E-2f;
  y = __fmaf_rn(y, x, -1.1514610310E-1f);
  y = __fmaf_rn(y, x, 1.1676998740E-1f);
  y = __fmaf_rn(y, x, -1.2420140846E-1f);
  y = __fmaf_rn(y, x, 1.4249322787E-1f);
  y = __fmaf_rn(y, x, -1.6668057665E-1f);
  y = __fmaf_rn(y, x, 2.0000714765E-1f);
  y = __fmaf_rn(y, x, -2.4999993993E-1f);
  y = __fmaf_rn(y, x, 3.3333331174E-1f);
  y = __fmul_rn(y, x);
  y = __fmaf_rn(y, z, __fmul_rn(e, -2.12194440e-4f));   // (y*x)*z fused with the add of e*q1
  y = __fsub_rn(y, __fmul_rn(z, 0.5f));                  // z/2 is exact
  x = __fadd_rn(x, y);
  return __fadd_rn(x, __fmul_rn(e, 0.693359375f));       // e*q2 is exact (q2 = 355/512)
}

__device__ __forceinline__ void cephes_sincosf_avx(float x, float& s, float& c) {
  uint32_t sign_sin = __float_as_uint(x) & 0x80000000u;
  x = fabsf(x);
  float y = __fmul_rn(x, 1.27323954473516f);
  int j = __float2int_rz(y);
  j = (j + 1) & ~1;
  y = __int2float_rn(j);
  sign_sin ^= ((uint32_t)(j & 4)) << 29;
  const bool poly = (j & 2) == 0;
  const uint32_t sign_cos = ((uint32_t)(~(j - 2) & 4)) << 29;
  x = __fmaf_rn(y, -0.78515625f, x);
  x = __fmaf_rn(y, -2.4187564849853515625e-4f, x);
  x = __fmaf_rn(y, -3.77489497744594108e-8f, x);
  const float z = __fmul_rn(x, x);
  float yc = 2.443315711809948E-005f;
  yc = __fmaf_rn(yc, z, -1.388731625493765E-003f);
  yc = __fmaf_rn(yc, z, 4.166664568298827E-002f);
  yc = __fmul_rn(yc, z);
  yc = __fmaf_rn(yc, z, -__fmul_rn(z, 0.5f));            // (yc*z)*z fused with the subtraction of z/2
  yc = __fadd_rn(yc, 1.0f);
  float ys = -1.9515295891E-4f;
  ys = __fmaf_rn(ys, z, 8.3321608736E-3f);
  ys = __fmaf_rn(ys, z, -1.6666654611E-1f);
  ys = __fmul_rn(ys, z);
  ys = __fmaf_rn(ys, x, x);
  s = __uint_as_float(__float_as_uint(poly ? ys : yc) ^ sign_sin);
  c = __uint_as_float(__float_as_uint(poly ? yc : ys) ^ sign_cos);
}

// One thread per Box-Muller pair (j, j+8) of a 16-group.  `src` holds the raw generator words of the group, `dst`
// receives the normals (src == dst for the body of the tensor; the redrawn tail reads the 16 extra words and writes the
// last 16 values).
__global__ void __launch_bounds__(kBlock) normal_fill16_kernel(const uint32_t* src, float* dst, int64_t n_pairs) {
  pdl_prologue();
  const int64_t p = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (p >= n_pairs) return;
  const int64_t i = (p >> 3) * 16 + (p & 7);
  const float u1 = __fsub_rn(1.0f, mt_uniform(src[i]));   // [0,1) -> (0,1]
  const float u2 = mt_uniform(src[i + 8]);
  const float radius = __fsqrt_rn(__fmul_rn(-2.0f, cephes_logf_avx(u1)));
  const float theta = __fmul_rn(6.2831854820251465f, u2);   // float(2 * pi<double>)
  float s, c;
  cephes_sincosf_avx(theta, s, c);
  dst[i] = __fmaf_rn(__fmul_rn(radius, c), 1.0f, 0.0f);      // fmadd(n, std = 1, mean = 0): -0 becomes +0
  dst[i + 8] = __fmaf_rn(__fmul_rn(radius, s), 1.0f, 0.0f);
}

}  // namespace lp

using namespace lp;

extern "C" int lp_torch_cpu_randn_f32(float* out, int64_t n, uint64_t seed, uint32_t* state_out, int64_t* consumed_out,
                                      lp_stream_t stream) {
  if (!out || n < 0) return LP_ERR_INVALID;
  if (n < 16 || n >= (int64_t(1) << 40)) return LP_ERR_UNSUPPORTED;   // below 16 torch takes another (scalar, double) path
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t tail = (n % 16) ? 16 : 0;          // the redrawn last 16 values consume 16 more outputs
  const int64_t consumed = n + tail;
  const int64_t n_words = ((consumed + kMtN - 1) / kMtN) * kMtN;   // through the end of the engine's current block
  if (consumed_out) *consumed_out = consumed;
  launch_kernel_ex(mt19937_raw_kernel, dim3(1), dim3(kMtThreads), 0, s, reinterpret_cast<uint32_t*>(out), consumed, n_words,
                   (uint32_t)(seed & 0xffffffffull), state_out);
  if (int rc = check_launch()) return rc;
  const int64_t pairs = (n / 16) * 8;
  launch_kernel(normal_fill16_kernel, dim3((unsigned)((pairs + kBlock - 1) / kBlock)), s,
                reinterpret_cast<const uint32_t*>(out), out, pairs);
  if (int rc = check_launch()) return rc;
  if (tail) {
    launch_kernel(normal_fill16_kernel, dim3(1), s, reinterpret_cast<const uint32_t*>(out + n), out + n - 16, (int64_t)8);
    if (int rc = check_launch()) return rc;
  }
  return LP_OK;
}
