// QDX-F32 arithmetic on the device: Threefry-2x32-20 (JAX partitionable mode), the uniform / normal
// bit tricks of jax.random, and the log1p / erfinv / sincos kernels -- every rounding step explicit.
// This translation unit family is compiled with -fmad=false, so a*b+c is two roundings unless
// __fmaf_rn is written.  The same operation sequence is restated (independently, in C) by
// oracle/qdx_oracle.c; tests compare the two bit for bit.
//
// Reference call sites (under /root/reference): jax.random.split/uniform/normal/choice in
// qdax/core/emitters/mutation_operators.py:205-220, repertoire_selectors/uniform_selector.py:48-55,
// qdax/core/map_elites.py:177-241; jnp.cos/sin/cumsum in qdax/tasks/arm.py:28-36 and
// qdax/tasks/standard_functions.py:13-15.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#define QDX_DEV __device__ __forceinline__

struct QdxKey { uint32_t a, b; };

QDX_DEV uint32_t qdx_rotl(uint32_t x, int r) { return __funnelshift_l(x, x, r); }

// Threefry-2x32, 20 rounds; key schedule ks = (k0, k1, k0^k1^0x1BD11BDA).
QDX_DEV void qdx_threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t& o0, uint32_t& o1) {
    const uint32_t k2 = k0 ^ k1 ^ 0x1BD11BDAu;
    uint32_t x0 = c0 + k0, x1 = c1 + k1;
#define QDX_R(r) x0 += x1; x1 = qdx_rotl(x1, r); x1 ^= x0;
    QDX_R(13) QDX_R(15) QDX_R(26) QDX_R(6)
    x0 += k1; x1 += k2 + 1u;
    QDX_R(17) QDX_R(29) QDX_R(16) QDX_R(24)
    x0 += k2; x1 += k0 + 2u;
    QDX_R(13) QDX_R(15) QDX_R(26) QDX_R(6)
    x0 += k0; x1 += k1 + 3u;
    QDX_R(17) QDX_R(29) QDX_R(16) QDX_R(24)
    x0 += k1; x1 += k2 + 4u;
    QDX_R(13) QDX_R(15) QDX_R(26) QDX_R(6)
    x0 += k2; x1 += k0 + 5u;
#undef QDX_R
    o0 = x0; o1 = x1;
}

// (Measured and rejected, profiles/r1_notes.md: issuing the rotations on the FMA pipe as IMAD.WIDE by 2^r + one LOP3
//  is 22 % slower on B200 -- the wide multiply is not full-rate -- so the funnel shift stays.)
// jax.random.split(key, n)[i]
QDX_DEV QdxKey qdx_split(QdxKey k, uint64_t i) {
    QdxKey o;
    qdx_threefry2x32(k.a, k.b, (uint32_t)(i >> 32), (uint32_t)i, o.a, o.b);
    return o;
}
// random_bits(key, 32, shape)[flat i]
QDX_DEV uint32_t qdx_bits32(QdxKey k, uint64_t i) {
    uint32_t a, b;
    qdx_threefry2x32(k.a, k.b, (uint32_t)(i >> 32), (uint32_t)i, a, b);
    return a ^ b;
}
QDX_DEV float qdx_unit_float(uint32_t bits) { return __uint_as_float((bits >> 9) | 0x3F800000u) - 1.0f; }

// jnp.maximum / jnp.minimum: NaN propagates like XLA max/min, -0 < +0 (IEEE 754-2019 maximum / minimum): FMNMX.NAN.
QDX_DEV float qdx_max_nanprop(float x, float lo) { float r; asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(lo)); return r; }
QDX_DEV float qdx_min_nanprop(float x, float hi) { float r; asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(hi)); return r; }

// log(t), t > 0 normal: t = m * 2^e, m in [sqrt(1/2), sqrt(2)).
QDX_DEV float qdx_logf(float t) {
    uint32_t ix = __float_as_uint(t) - 0x3f3504f3u;
    int e = (int32_t)ix >> 23;
    float m = __uint_as_float((ix & 0x007fffffu) + 0x3f3504f3u);
    float r = m - 1.0f;
    float z = r * r;
    float p = 7.0376836292E-2f;
    p = __fmaf_rn(p, r, -1.1514610310E-1f);
    p = __fmaf_rn(p, r, 1.1676998740E-1f);
    p = __fmaf_rn(p, r, -1.2420140846E-1f);
    p = __fmaf_rn(p, r, 1.4249322787E-1f);
    p = __fmaf_rn(p, r, -1.6668057665E-1f);
    p = __fmaf_rn(p, r, 2.0000714765E-1f);
    p = __fmaf_rn(p, r, -2.4999993993E-1f);
    p = __fmaf_rn(p, r, 3.3333331174E-1f);
    float y = (p * r) * z;
    float fe = (float)e;
    y = __fmaf_rn(fe, -2.12194440e-4f, y);
    y = __fmaf_rn(z, -0.5f, y);
    float res = r + y;
    res = __fmaf_rn(fe, 0.693359375f, res);
    return res;
}
// log1p(y), y in (-1, 0]: log(t) + (y - (t - 1)) * (2 - t), t = fl(1 + y); branch- and division-free
QDX_DEV float qdx_log1pf(float y) {
    float t = 1.0f + y;
    float c = (y - (t - 1.0f)) * (2.0f - t);
    return qdx_logf(t) + c;
}
// XLA ErfInv32 (Giles), fused Horner.  EDGE: handle |x| == 1 (never produced by jax.random.normal's uniform).
// The tail polynomial (w >= 5, 0.3 % of draws) is entered by a warp-uniform vote so the common path is straight-line.
// FULLWARP: the caller guarantees all 32 lanes are converged here (saves the activemask query).
QDX_DEV float qdx_erfinv_central_poly(float w) {
    w = w - 2.5f;
    float p = 2.81022636e-08f;
    p = __fmaf_rn(p, w, 3.43273939e-07f);
    p = __fmaf_rn(p, w, -3.5233877e-06f);
    p = __fmaf_rn(p, w, -4.39150654e-06f);
    p = __fmaf_rn(p, w, 0.00021858087f);
    p = __fmaf_rn(p, w, -0.00125372503f);
    p = __fmaf_rn(p, w, -0.00417768164f);
    p = __fmaf_rn(p, w, 0.246640727f);
    p = __fmaf_rn(p, w, 1.50140941f);
    return p;
}
// p(w) given w = -log1p(-x^2): central polynomial when the whole warp is central, else per-lane coefficients
template <bool FULLWARP>
QDX_DEV float qdx_erfinv_poly(float w) {
    float p;
    const bool central = w < 5.0f;
    if (__all_sync(FULLWARP ? 0xffffffffu : __activemask(), central)) {
        p = qdx_erfinv_central_poly(w);
    } else {
        // mixed warp: both polynomials share one Horner chain with per-lane coefficients
        w = central ? w - 2.5f : __fsqrt_rn(w) - 3.0f;
        p = central ? 2.81022636e-08f : -0.000200214257f;
        p = __fmaf_rn(p, w, central ? 3.43273939e-07f : 0.000100950558f);
        p = __fmaf_rn(p, w, central ? -3.5233877e-06f : 0.00134934322f);
        p = __fmaf_rn(p, w, central ? -4.39150654e-06f : -0.00367342844f);
        p = __fmaf_rn(p, w, central ? 0.00021858087f : 0.00573950773f);
        p = __fmaf_rn(p, w, central ? -0.00125372503f : -0.0076224613f);
        p = __fmaf_rn(p, w, central ? -0.00417768164f : 0.00943887047f);
        p = __fmaf_rn(p, w, central ? 0.246640727f : 1.00167406f);
        p = __fmaf_rn(p, w, central ? 1.50140941f : 2.83297682f);
    }
    return p;
}
template <bool EDGE, bool FULLWARP>
QDX_DEV float qdx_erfinvf_t(float x) {
    const float w = -qdx_log1pf(-(x * x));
    const float p = qdx_erfinv_poly<FULLWARP>(w);
    if (EDGE && fabsf(x) == 1.0f) return x * 3.40282347e+38f;
    return p * x;
}
QDX_DEV float qdx_erfinvf(float x) { return qdx_erfinvf_t<true, false>(x); }
// jax.random.normal from one 32-bit draw.
template <bool FULLWARP>
QDX_DEV float qdx_normal_from_bits_t(uint32_t bits) {
    const float lo = -0x1.fffffep-1f;
    float f = qdx_unit_float(bits);
    float u = f * 2.0f + lo;
    u = u < lo ? lo : u;
    return 0x1.6a09e6p+0f * qdx_erfinvf_t<false, FULLWARP>(u);   // u in [lo, 1): |u| == 1 impossible
}
QDX_DEV float qdx_normal_from_bits(uint32_t bits) { return qdx_normal_from_bits_t<false>(bits); }
// ---- packed FP32 pairs (Blackwell FADD2 / FMUL2 / FFMA2, PTX add / sub / mul / fma .rn.f32x2): two IEEE round-to-nearest
// results per issued instruction.  Each element sees exactly the operation the scalar code performs, so the results are
// bit-identical; the kernels that use them are bound by instruction issue, not by the FMA pipe.  (-DQDX_PACKED_F32=0: scalar.)
#ifndef QDX_PACKED_F32
#define QDX_PACKED_F32 1
#endif
struct QdxF2 { unsigned long long v; };
QDX_DEV QdxF2 qdx_f2(float a, float b) { QdxF2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(a), "f"(b)); return r; }
QDX_DEV QdxF2 qdx_f2(float a) { return qdx_f2(a, a); }
QDX_DEV void qdx_f2_get(QdxF2 x, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(x.v)); }
QDX_DEV QdxF2 operator+(QdxF2 a, QdxF2 b) { QdxF2 r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
QDX_DEV QdxF2 operator-(QdxF2 a, QdxF2 b) { QdxF2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
QDX_DEV QdxF2 operator*(QdxF2 a, QdxF2 b) { QdxF2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r.v) : "l"(a.v), "l"(b.v)); return r; }
QDX_DEV QdxF2 qdx_fma2(QdxF2 a, QdxF2 b, QdxF2 c) { QdxF2 r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r.v) : "l"(a.v), "l"(b.v), "l"(c.v)); return r; }
// qdx_logf on a pair (the integer decomposition stays scalar)
QDX_DEV QdxF2 qdx_logf2(float t0, float t1) {
    const uint32_t i0 = __float_as_uint(t0) - 0x3f3504f3u, i1 = __float_as_uint(t1) - 0x3f3504f3u;
    const int e0 = (int32_t)i0 >> 23, e1 = (int32_t)i1 >> 23;
    const QdxF2 m = qdx_f2(__uint_as_float((i0 & 0x007fffffu) + 0x3f3504f3u), __uint_as_float((i1 & 0x007fffffu) + 0x3f3504f3u));
    const QdxF2 r = m - qdx_f2(1.0f);
    const QdxF2 z = r * r;
    QdxF2 p = qdx_f2(7.0376836292E-2f);
    p = qdx_fma2(p, r, qdx_f2(-1.1514610310E-1f));
    p = qdx_fma2(p, r, qdx_f2(1.1676998740E-1f));
    p = qdx_fma2(p, r, qdx_f2(-1.2420140846E-1f));
    p = qdx_fma2(p, r, qdx_f2(1.4249322787E-1f));
    p = qdx_fma2(p, r, qdx_f2(-1.6668057665E-1f));
    p = qdx_fma2(p, r, qdx_f2(2.0000714765E-1f));
    p = qdx_fma2(p, r, qdx_f2(-2.4999993993E-1f));
    p = qdx_fma2(p, r, qdx_f2(3.3333331174E-1f));
    QdxF2 y = (p * r) * z;
    const QdxF2 fe = qdx_f2((float)e0, (float)e1);
    y = qdx_fma2(fe, qdx_f2(-2.12194440e-4f), y);
    y = qdx_fma2(z, qdx_f2(-0.5f), y);
    QdxF2 res = r + y;
    res = qdx_fma2(fe, qdx_f2(0.693359375f), res);
    return res;
}
// qdx_log1pf on a pair.  ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into FFMA2 (it does not for the scalar .rn forms), which
// changes the rounding: wherever the spec has a product feeding a sum, the product (or the sum) is done in scalar instructions.
QDX_DEV void qdx_log1pf2(float y0, float y1, float& l0, float& l1) {
    const QdxF2 y = qdx_f2(y0, y1);
    const QdxF2 t = qdx_f2(1.0f) + y;
    const QdxF2 a = y - (t - qdx_f2(1.0f)), b = qdx_f2(2.0f) - t;
    float t0, t1, a0, a1, b0, b1, g0, g1;
    qdx_f2_get(t, t0, t1); qdx_f2_get(a, a0, a1); qdx_f2_get(b, b0, b1);
    qdx_f2_get(qdx_logf2(t0, t1), g0, g1);
    const float c0 = a0 * b0, c1 = a1 * b1;
    l0 = g0 + c0; l1 = g1 + c1;
}
// qdx_erfinv_central_poly on a pair
QDX_DEV QdxF2 qdx_erfinv_central_poly2(QdxF2 w) {
    w = w - qdx_f2(2.5f);
    QdxF2 p = qdx_f2(2.81022636e-08f);
    p = qdx_fma2(p, w, qdx_f2(3.43273939e-07f));
    p = qdx_fma2(p, w, qdx_f2(-3.5233877e-06f));
    p = qdx_fma2(p, w, qdx_f2(-4.39150654e-06f));
    p = qdx_fma2(p, w, qdx_f2(0.00021858087f));
    p = qdx_fma2(p, w, qdx_f2(-0.00125372503f));
    p = qdx_fma2(p, w, qdx_f2(-0.00417768164f));
    p = qdx_fma2(p, w, qdx_f2(0.246640727f));
    p = qdx_fma2(p, w, qdx_f2(1.50140941f));
    return p;
}

// (Measured, not kept: evaluating BOTH erfinv polynomials on packed pairs in mixed warps and selecting per element -- bit-identical,
//  c3 0.6465 -> 0.654 ms: the per-lane coefficient selects of the scalar mixed path are cheaper than a second polynomial.)
// Four normals from four draws, all 32 lanes converged: the Threefry blocks that produced `bits`, the uniform -> w
// transforms and (when all 128 draws of the warp are central, 68 % of the time) the four Horner chains are straight-line
// code with four independent dependency chains -- the per-draw version is one serial chain with a branch per draw.
// Same operations per draw, so the results are bit-identical to qdx_normal_from_bits.
QDX_DEV void qdx_normal4_from_bits(const uint32_t bits[4], float out[4]) {
    const float lo = -0x1.fffffep-1f;
#if QDX_PACKED_F32
    float u[4], w[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const QdxF2 f = qdx_f2(qdx_unit_float(bits[2 * h]), qdx_unit_float(bits[2 * h + 1]));
        const QdxF2 v = f * qdx_f2(2.0f) + qdx_f2(lo);
        float v0, v1; qdx_f2_get(v, v0, v1);
        u[2 * h] = v0 < lo ? lo : v0; u[2 * h + 1] = v1 < lo ? lo : v1;
        float l0, l1;
        qdx_log1pf2(-(u[2 * h] * u[2 * h]), -(u[2 * h + 1] * u[2 * h + 1]), l0, l1);     // the squares in scalar: see qdx_log1pf2
        w[2 * h] = -l0; w[2 * h + 1] = -l1;
    }
    const bool central2 = (w[0] < 5.0f) & (w[1] < 5.0f) & (w[2] < 5.0f) & (w[3] < 5.0f);
    if (__all_sync(0xffffffffu, central2)) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const QdxF2 r = qdx_f2(0x1.6a09e6p+0f) * (qdx_erfinv_central_poly2(qdx_f2(w[2 * h], w[2 * h + 1])) * qdx_f2(u[2 * h], u[2 * h + 1]));
            qdx_f2_get(r, out[2 * h], out[2 * h + 1]);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = 0x1.6a09e6p+0f * (qdx_erfinv_poly<true>(w[j]) * u[j]);
    }
#else
    float u[4], w[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float f = qdx_unit_float(bits[j]);
        float v = f * 2.0f + lo;
        u[j] = v < lo ? lo : v;
        w[j] = -qdx_log1pf(-(u[j] * u[j]));
    }
    const bool central = (w[0] < 5.0f) & (w[1] < 5.0f) & (w[2] < 5.0f) & (w[3] < 5.0f);
    if (__all_sync(0xffffffffu, central)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = 0x1.6a09e6p+0f * (qdx_erfinv_central_poly(w[j]) * u[j]);
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) out[j] = 0x1.6a09e6p+0f * (qdx_erfinv_poly<true>(w[j]) * u[j]);
    }
#endif
}
// sin & cos: 3-term Cody-Waite by pi/2 (fused), minimax kernels on [-pi/4, pi/4].
QDX_DEV void qdx_sincosf(float th, float& s_out, float& c_out) {
    float q = rintf(th * 0x1.45f306p-1f);
    float r = __fmaf_rn(q, -0x1.921fb6p+0f, th);
    r = __fmaf_rn(q, 0x1.777a5cp-25f, r);
    r = __fmaf_rn(q, 0x1.ee59dap-50f, r);
    int n = (int)q & 3;
    float s = r * r;
    float ps = __fmaf_rn(__fmaf_rn(-1.9515295891E-4f, s, 8.3321608736E-3f), s, -1.6666654611E-1f);
    float sr = __fmaf_rn(r * s, ps, r);
    float pc = __fmaf_rn(__fmaf_rn(2.443315711809948E-5f, s, -1.388731625493765E-3f), s, 4.166664568298827E-2f);
    float cr = __fmaf_rn(s * s, pc, __fmaf_rn(s, -0.5f, 1.0f));
    float sv = (n & 1) ? cr : sr;
    float cv = (n & 1) ? sr : cr;
    if (n & 2) sv = -sv;
    if ((n + 1) & 2) cv = -cv;
    s_out = sv; c_out = cv;
}

// qdx_sincosf on a pair: the reduction and the two minimax kernels on packed FP32 (every sum sits inside an FMA, so nothing here
// can be contracted differently from the scalar code); rounding to the quadrant, the quadrant selects and the signs stay scalar.
QDX_DEV void qdx_sincosf2(float th0, float th1, float& s0, float& c0, float& s1, float& c1) {
#if QDX_PACKED_F32
    const float q0 = rintf(th0 * 0x1.45f306p-1f), q1 = rintf(th1 * 0x1.45f306p-1f);
    const QdxF2 q = qdx_f2(q0, q1);
    QdxF2 r = qdx_fma2(q, qdx_f2(-0x1.921fb6p+0f), qdx_f2(th0, th1));
    r = qdx_fma2(q, qdx_f2(0x1.777a5cp-25f), r);
    r = qdx_fma2(q, qdx_f2(0x1.ee59dap-50f), r);
    const QdxF2 s = r * r;
    const QdxF2 ps = qdx_fma2(qdx_fma2(qdx_f2(-1.9515295891E-4f), s, qdx_f2(8.3321608736E-3f)), s, qdx_f2(-1.6666654611E-1f));
    const QdxF2 sr = qdx_fma2(r * s, ps, r);
    const QdxF2 pc = qdx_fma2(qdx_fma2(qdx_f2(2.443315711809948E-5f), s, qdx_f2(-1.388731625493765E-3f)), s, qdx_f2(4.166664568298827E-2f));
    const QdxF2 cr = qdx_fma2(s * s, pc, qdx_fma2(s, qdx_f2(-0.5f), qdx_f2(1.0f)));
    float sr0, sr1, cr0, cr1;
    qdx_f2_get(sr, sr0, sr1); qdx_f2_get(cr, cr0, cr1);
    const int n0 = (int)q0 & 3, n1 = (int)q1 & 3;
    float sv0 = (n0 & 1) ? cr0 : sr0, cv0 = (n0 & 1) ? sr0 : cr0;
    float sv1 = (n1 & 1) ? cr1 : sr1, cv1 = (n1 & 1) ? sr1 : cr1;
    if (n0 & 2) sv0 = -sv0;
    if ((n0 + 1) & 2) cv0 = -cv0;
    if (n1 & 2) sv1 = -sv1;
    if ((n1 + 1) & 2) cv1 = -cv1;
    s0 = sv0; c0 = cv0; s1 = sv1; c1 = cv1;
#else
    qdx_sincosf(th0, s0, c0); qdx_sincosf(th1, s1, c1);
#endif
}

// exp(z): n = rint(z*log2e), two-term fused reduction by ln2, degree-6 polynomial, two-step scaling by 2^n
QDX_DEV float qdx_expf(float z) {
    if (z != z) return z;
    if (z > 88.75f) return INFINITY;
    if (z < -104.0f) return 0.0f;
    float n = rintf(z * 0x1.715476p+0f);
    float r = __fmaf_rn(n, -0x1.62e400p-1f, z);
    r = __fmaf_rn(n, -0x1.7f7d1cp-20f, r);
    float p = 0x1.6c16c2p-10f;
    p = __fmaf_rn(p, r, 0x1.111112p-7f);
    p = __fmaf_rn(p, r, 0x1.555556p-5f);
    p = __fmaf_rn(p, r, 0x1.555556p-3f);
    p = __fmaf_rn(p, r, 0.5f);
    p = __fmaf_rn(p, r, 1.0f);
    p = __fmaf_rn(p, r, 1.0f);
    int ni = (int)n;
    int h = ni / 2;
    float s1 = __uint_as_float((uint32_t)(h + 127) << 23), s2 = __uint_as_float((uint32_t)(ni - h + 127) << 23);
    return (p * s1) * s2;
}
// pow(x, y) = exp(y * log(x)) for x >= 0 (polynomial mutation); pow(0, y > 0) = 0; x < 0 -> NaN
QDX_DEV float qdx_powf(float x, float y) {
    if (x != x || y != y) return __int_as_float(0x7fc00000);
    if (x < 0.0f) return __int_as_float(0x7fc00000);
    if (x == 0.0f) return y > 0.0f ? 0.0f : (y == 0.0f ? 1.0f : INFINITY);
    if (x == INFINITY) return y > 0.0f ? INFINITY : (y == 0.0f ? 1.0f : 0.0f);
    float lg;
    if (x < 0x1p-126f) lg = qdx_logf(x * 0x1p+25f) - 0x1.154246p+4f;
    else lg = qdx_logf(x);
    return qdx_expf(y * lg);
}

// Total order key for a fitness: -inf < ... < -0 == +0 < ... < +inf < NaN (NaN = 0xFFFFFFFF).
QDX_DEV uint32_t qdx_order_key(float v) {
    if (v != v) return 0xFFFFFFFFu;
    uint32_t u = __float_as_uint(v == 0.0f ? 0.0f : v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
