// Uniform parent selection: jax.random.choice(key, arange(K), (B,), p=occupied/M, replace=True)
// (qdax/core/emitters/repertoire_selectors/uniform_selector.py:43-55 under /root/reference).
//
//   cum = cumsum(p)            sequential float32 running sum (canonical order)
//   r   = cum[-1] * (1 - u)    u = uniform(key, (B,))
//   idx = searchsorted(cum, r, side='left')
//
// All non-zero p are the same value q = fl(1/M), so cum over the K cells is a step function of the
// number j of occupied cells seen so far:  cum = T[j],  T[j] = fl(T[j-1] + q),  T[0] = 0, and the selected
// cell is the j-th occupied cell with j = min{ j : T[j] >= r }.  T is NOT materialised: inside one binade
// a running float32 sum advances by a constant step once its rounding parity has settled, so T is a short
// list of arithmetic segments (<= 3 per binade).  qdx_build_sel constructs the segments with real float32
// additions at every irregular step, so T is reproduced bit for bit (tests/test_host_logic.py checks every
// M up to 70 000 and samples up to 4 M against np.cumsum).
#pragma once
#include <cstdint>
#include <cmath>

#ifdef __CUDACC__
#define QDX_HD __host__ __device__ inline
#else
#define QDX_HD inline
#endif

#define QDX_MAX_SEG 128

struct QdxSeg {
    int32_t j0;     // first rank covered by this segment (1-based)
    int32_t n;      // T[j0 + i] = s0 + i * delta for i in [0, n]
    float s0;
    float delta;
};

struct QdxSel {
    int32_t M;        // number of occupied cells
    int32_t nseg;
    float total;      // T[M] = cum[-1]
    float q;          // fl(1 / M)
    QdxSeg seg[QDX_MAX_SEG];
    float last[QDX_MAX_SEG];   // last[s] = T[j0 + n] of segment s
};

QDX_HD uint32_t qdx_f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; return c.u;
#endif
}
QDX_HD float qdx_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}
// x + y with exactly one IEEE rounding on both host and device (no contraction possible: single op).
QDX_HD float qdx_add_rn(float x, float y) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(x, y);
#else
    volatile float r = x + y; return r;
#endif
}
QDX_HD float qdx_seg_value(const QdxSeg& s, int32_t i) {
    // exact: s0 + i*delta is representable for 0 <= i <= n
#ifdef __CUDA_ARCH__
    return __fmaf_rn((float)i, s.delta, s.s0);
#else
    return fmaf((float)i, s.delta, s.s0);
#endif
}

QDX_HD void qdx_build_sel(int32_t M, QdxSel* out) {
    out->M = M; out->nseg = 0; out->total = 0.0f; out->q = 0.0f;
    if (M <= 0) return;
    const float q = 1.0f / (float)M;
    out->q = q;
    int ns = 0;
    int32_t j = 1;
    float prev = q;                       // T[1]
    out->seg[ns] = QdxSeg{1, 0, q, 0.0f}; out->last[ns] = q; ++ns;
    j = 2;
    while (j <= M && ns < QDX_MAX_SEG) {
        float V = qdx_add_rn(prev, q);    // T[j]
        const uint32_t vb = qdx_f2u(V);
        bool same_binade = (vb >> 23) == (qdx_f2u(prev) >> 23);
        int32_t n = 0; float delta = 0.0f;
        if (same_binade && j < M) {
            float W = qdx_add_rn(V, q);   // T[j+1]
            const uint32_t topb = vb | 0x007fffffu;           // largest float of V's binade
            if (qdx_f2u(W) <= topb) {                         // positive floats order like their bits
                delta = W - V;                                 // exact
                // integer mantissa units of V's binade: ulp = 2^(e-23)
                const uint32_t units_left = topb - vb;         // (top - V) / ulp
                const int eV = (int)(vb >> 23);
                // delta / ulp: delta is a multiple of ulp and < 2^24 ulp
                const uint32_t db = qdx_f2u(delta);
                const int eD = (int)(db >> 23);
                const uint32_t mant = (db & 0x007fffffu) | 0x00800000u;   // delta = mant * 2^(eD-150)
                const int sh = eV - eD;                                    // ulp(V) = 2^(eV-150)
                const uint32_t dunits = sh >= 0 ? (sh < 32 ? (mant >> sh) : 0u) : (mant << (-sh));
                if (dunits > 0) {
                    n = (int32_t)(units_left / dunits);
                    if (n > M - j) n = M - j;
                }
            }
        }
        QdxSeg s{j, n, V, delta};
        const float lastv = n > 0 ? qdx_seg_value(s, n) : V;
        out->seg[ns] = s; out->last[ns] = lastv; ++ns;
        prev = lastv;
        j += n + 1;
    }
    out->nseg = ns;
    out->total = out->last[ns - 1];
    if (j <= M) out->nseg = -1;   // segment table overflow: caller must treat as an error
}

// rank j in [1, M] with T[j] >= r (first such); r in (0, total].
QDX_HD int32_t qdx_sel_rank(const QdxSeg* seg, const float* last, int32_t nseg, float r) {
    int s = nseg - 1;
    while (s > 0 && last[s - 1] >= r) --s;
    const QdxSeg sg = seg[s];
    if (sg.n == 0 || r <= sg.s0) return sg.j0;
    int32_t i = (int32_t)((r - sg.s0) / sg.delta);
    if (i > sg.n) i = sg.n;
    while (i > 0 && qdx_seg_value(sg, i - 1) >= r) --i;
    while (i < sg.n && qdx_seg_value(sg, i) < r) ++i;
    return sg.j0 + i;
}
