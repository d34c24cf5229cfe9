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
// consecutive words depends on the same column of the previous row (same thread) and on two words 2.7 rows back
// (other threads): ONE CTA walks the stream with 227 threads, the last 2048 words in a shared-memory ring, one
// barrier per TWO rows.  The transform is a separate, fully parallel launch.
#include "lp_common.cuh"

namespace lp {

constexpr int kMtN = 624, kMtM = 397, kMtLag = kMtN - kMtM;  // 227
constexpr int kMtRing = 2048;                                  // >= 624 + 2*227 + slack, power of two
constexpr int kMtThreads = 256;

__device__ __forceinline__ uint32_t mt_twist(uint32_t u, uint32_t v) {
  const uint32_t y = (u & 0x80000000u) | (v & 0x7fffffffu);
  return (y >> 1) ^ ((v & 1u) ? 0x9908b0dfu : 0u);
}

__device__ __forceinline__ float mt_uniform(uint32_t y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return __uint2float_rn(y & 0x00ffffffu) * 5.9604644775390625e-08f;  // exact: 24 bits times 2^-24
}

// out[i] = uniform of the i-th generator output for i < n_out; words are generated up to n_words (a multiple of 624,
// >= n_out) so that state_out receives the engine's complete state array after its last twist.
__global__ void __launch_bounds__(kMtThreads, 1) mt19937_uniform_kernel(float* __restrict__ out, int64_t n_out,
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
  // x[kMtN + i] is generator output i (before tempering).  Row r covers outputs [227 r, 227 r + 227).
  const int64_t rows = (n_words + kMtLag - 1) / kMtLag;
  for (int64_t r = 0; r < rows; r += 2) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t i = (r + h) * kMtLag + c;   // output index of this thread in this row
      if (c < kMtLag && i < n_words) {
        const uint32_t n = (uint32_t)((i + kMtN) & (kMtRing - 1));
        const uint32_t a = ring[(n - kMtN) & (kMtRing - 1)];
        const uint32_t b = ring[(n - kMtN + 1) & (kMtRing - 1)];
        const uint32_t m = ring[(n - kMtLag) & (kMtRing - 1)];
        const uint32_t x = m ^ mt_twist(a, b);
        ring[n] = x;
        if (i < n_out) out[i] = mt_uniform(x);
        if (state_out && i >= n_words - kMtN) state_out[i - (n_words - kMtN)] = x;
      }
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

// One thread per Box-Muller pair (j, j+8) of a 16-group.  `src` holds the uniforms of the groups, `dst` receives the
// normals (src == dst for the body of the tensor; the redrawn tail reads the 16 extra uniforms and writes the last 16).
__global__ void __launch_bounds__(kBlock) normal_fill16_kernel(const float* src, float* dst, int64_t n_pairs) {
  pdl_prologue();
  const int64_t p = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (p >= n_pairs) return;
  const int64_t i = (p >> 3) * 16 + (p & 7);
  const float u1 = __fsub_rn(1.0f, src[i]);   // [0,1) -> (0,1]
  const float u2 = src[i + 8];
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
  launch_kernel_ex(mt19937_uniform_kernel, dim3(1), dim3(kMtThreads), 0, s, out, consumed, n_words,
                   (uint32_t)(seed & 0xffffffffull), state_out);
  if (int rc = check_launch()) return rc;
  const int64_t pairs = (n / 16) * 8;
  launch_kernel(normal_fill16_kernel, dim3((unsigned)((pairs + kBlock - 1) / kBlock)), s, (const float*)out, out, pairs);
  if (int rc = check_launch()) return rc;
  if (tail) {
    launch_kernel(normal_fill16_kernel, dim3(1), s, (const float*)(out + n), out + n - 16, (int64_t)8);
    if (int rc = check_launch()) return rc;
  }
  return LP_OK;
}
