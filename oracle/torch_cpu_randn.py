"""CPU restatement of `torch.randn(n, generator=torch.Generator().manual_seed(seed), device="cpu")` (fp32, n >= 16).

TEST INFRASTRUCTURE ONLY (nothing under lanpaint_b200/ imports it).  It is the written-down form of what
`lanpaint_b200/csrc/lp_hostnoise.cu` computes on the device, in numpy, operation for operation, so that the algorithm
and -- above all -- WHICH multiply-adds the host compiler fused are pinned by a CPU test against torch itself
(tests/test_oracle_torch_cpu_randn.py), independently of a GPU:

  * at::mt19937 (ATen/core/MT19937RNGEngine.h): init_genrand(seed & 0xffffffff), x[n] = x[n-227] ^ twist(x[n-624], x[n-623]),
    tempering; one 32-bit output per value, u = (y & 0xffffff) * 2^-24 (ATen/core/TransformationHelper.h uniform_real);
  * normal_fill (ATen/native/cpu/DistributionTemplates.h): Box-Muller over the pairs (j, j+8) of every 16 values,
    u1 = 1 - u_j, radius = sqrt(-2 log u1), theta = float(2 pi) * u_{j+8}; a size that is not a multiple of 16 redraws
    its last 16 values from 16 fresh outputs;
  * log / sincos: avx_mathfun.h's log256_ps / sincos256_ps (single-precision cephes), as the AVX2 translation unit of
    torch executes them (the AVX512 dispatch has no kernel of its own for this op and falls back to it): compiled with
    FMA contraction, every Horner step and range-reduction step fused, and at the two places where an expression
    offers the compiler a choice --  y*z + e*q1  in log,  y*z - z/2  in cos -- the LAST multiply is the fused one.
    Found by comparing the variants with torch.randn: 0 mismatches in 5 M values for this one.

fp32 FMA is emulated as float32(longdouble(a) * longdouble(b) + longdouble(c)): the product is exact and the sum is
rounded to 64 bits before the rounding to 24, so a double-rounding difference needs a 2^-40 coincidence.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32
_LD = np.longdouble


def init_genrand(seed: int) -> np.ndarray:
    mt = np.zeros(624, dtype=np.uint64)
    mt[0] = seed & 0xFFFFFFFF
    for i in range(1, 624):
        mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> np.uint64(30))) + i) & 0xFFFFFFFF
    return mt.astype(np.uint32)


def mt19937_words(seed: int, n: int) -> np.ndarray:
    """the first n tempered outputs; the recurrence is walked in rows of 227 words like the device kernel does"""
    x = np.zeros(624 + ((n + 226) // 227) * 227, dtype=np.uint32)
    x[:624] = init_genrand(seed)
    for r in range((n + 226) // 227):
        X = 624 + 227 * r
        a, b, m = x[X - 624:X - 624 + 227], x[X - 623:X - 623 + 227], x[X - 227:X]
        # columns >= 170 of b / a reach into this very row only for X-623+226 = X-397 < X: never; all operands are complete
        y = (a & np.uint32(0x80000000)) | (b & np.uint32(0x7FFFFFFF))
        x[X:X + 227] = m ^ (y >> np.uint32(1)) ^ np.where(b & np.uint32(1), np.uint32(0x9908B0DF), np.uint32(0))
    y = x[624:624 + n].copy()
    y ^= y >> np.uint32(11)
    y ^= (y << np.uint32(7)) & np.uint32(0x9D2C5680)
    y ^= (y << np.uint32(15)) & np.uint32(0xEFC60000)
    y ^= y >> np.uint32(18)
    return y


def uniforms(seed: int, n: int) -> np.ndarray:
    return (mt19937_words(seed, n) & np.uint32(0xFFFFFF)).astype(f32) * f32(2.0 ** -24)


def _fma(a, b, c):
    return (np.asarray(a, dtype=_LD) * np.asarray(b, dtype=_LD) + np.asarray(c, dtype=_LD)).astype(f32)


def _c(v, like):
    return np.full_like(like, f32(v))


def cephes_logf_avx(x: np.ndarray) -> np.ndarray:
    x = np.maximum(x, np.array([0x00800000], dtype=np.uint32).view(f32)[0])
    bits = x.view(np.uint32)
    e = ((bits >> np.uint32(23)).astype(np.int32) - 0x7F).astype(f32) + f32(1)
    x = ((bits & np.uint32(0x807FFFFF)) | np.uint32(0x3F000000)).view(f32)
    lt = x < f32(0.707106781186547524)
    tmp = np.where(lt, x, f32(0))
    x = (x - f32(1)).astype(f32)
    e = (e - np.where(lt, f32(1), f32(0))).astype(f32)
    x = (x + tmp).astype(f32)
    z = (x * x).astype(f32)
    y = _c(7.0376836292E-2, x)
    for p in (-1.1514610310E-1, 1.1676998740E-1, -1.2420140846E-1, 1.4249322787E-1, -1.6668057665E-1, 2.0000714765E-1,
              -2.4999993993E-1, 3.3333331174E-1):
        y = _fma(y, x, _c(p, x))
    y = (y * x).astype(f32)
    y = _fma(y, z, (e * f32(-2.12194440e-4)).astype(f32))     # (y*x)*z fused with the add of e*q1
    y = (y - (z * f32(0.5)).astype(f32)).astype(f32)          # z/2 is exact
    x = (x + y).astype(f32)
    return (x + (e * f32(0.693359375)).astype(f32)).astype(f32)   # e*q2 is exact


def cephes_sincosf_avx(x: np.ndarray):
    neg = np.signbit(x)
    x = np.abs(x)
    y = (x * f32(1.27323954473516)).astype(f32)
    j = y.astype(np.int32)                  # truncation, like cvttps
    j = (j + 1) & ~1
    y = j.astype(f32)
    sign_sin = neg ^ ((j & 4) != 0)
    poly = (j & 2) == 0
    sign_cos = ((~(j - 2)) & 4) != 0
    for dp in (-0.78515625, -2.4187564849853515625e-4, -3.77489497744594108e-8):
        x = _fma(y, _c(dp, y), x)
    z = (x * x).astype(f32)
    yc = _c(2.443315711809948E-005, x)
    yc = _fma(yc, z, _c(-1.388731625493765E-003, x))
    yc = _fma(yc, z, _c(4.166664568298827E-002, x))
    yc = (yc * z).astype(f32)
    yc = _fma(yc, z, -(z * f32(0.5)).astype(f32))             # (yc*z)*z fused with the subtraction of z/2
    yc = (yc + f32(1)).astype(f32)
    ys = _c(-1.9515295891E-4, x)
    ys = _fma(ys, z, _c(8.3321608736E-3, x))
    ys = _fma(ys, z, _c(-1.6666654611E-1, x))
    ys = (ys * z).astype(f32)
    ys = _fma(ys, x, x)
    s, c = np.where(poly, ys, yc), np.where(poly, yc, ys)
    return np.where(sign_sin, -s, s).astype(f32), np.where(sign_cos, -c, c).astype(f32)


def _normal_fill16(u: np.ndarray) -> np.ndarray:
    u = u.reshape(-1, 16)
    u1 = (f32(1) - u[:, :8]).astype(f32)
    radius = np.sqrt((f32(-2) * cephes_logf_avx(u1)).astype(f32)).astype(f32)
    theta = (f32(2.0 * np.pi) * u[:, 8:]).astype(f32)
    s, c = cephes_sincosf_avx(theta)
    out = np.concatenate([(radius * c).astype(f32), (radius * s).astype(f32)], axis=1)
    return (out + f32(0)).astype(f32).reshape(-1)             # fmadd(n, std = 1, mean = 0): -0 becomes +0


def torch_cpu_randn(seed: int, n: int) -> np.ndarray:
    """-> (values [n] fp32, number of generator outputs consumed)"""
    if n < 16:
        raise ValueError("torch takes another (scalar, double) path below 16 values")
    consumed = n + (16 if n % 16 else 0)
    u = uniforms(seed, consumed)
    out = np.empty(n, dtype=f32)
    body = (n // 16) * 16
    out[:body] = _normal_fill16(u[:body])
    if n % 16:
        out[n - 16:] = _normal_fill16(u[n:n + 16])           # "recompute the last 16 values"
    return out, consumed
