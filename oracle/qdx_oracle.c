/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product
 * path (qdax_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load liboracle.so.
 *
 * Plain-C restatement of the MAP-Elites generation step of QDax 0.5.1 (reference mounted
 * read-only at /root/reference; file:line citations below are into that tree), with EVERY
 * floating-point rounding step spelled out ("QDX-F32 arithmetic spec"):
 *   - all arithmetic is IEEE-754 binary32, round-to-nearest-even, one rounding per written
 *     operation; fmaf() is used exactly where written; the file must be compiled with
 *     -ffp-contract=off and without fast-math;
 *   - reductions over the genotype axis are sequential, left to right, starting from +0.0f (XLA reduce init);
 *     squared distances over the descriptor axis start from the first term;
 *   - log1p / sin / cos are the polynomial kernels written in this file (the reference calls
 *     jnp.log1p / jnp.sin / jnp.cos whose last-bit behaviour belongs to jaxlib, which is not
 *     installable here); they agree with libm to ~1 ulp (tests/test_oracle_cross.py).
 * The CUDA kernels follow the same spec, which is what makes whole runs bit-comparable.
 *
 * The PRNG is jax==0.8.0's Threefry-2x32 in partitionable mode (third-party, un-vendored,
 * uv.lock:1128-1129): restated from the published algorithm, pinned by the Random123 KATs
 * and by three values recalled from the JAX documentation (tests/test_oracle_prng.py).
 * PARITY UNPINNED at the jaxlib boundary for everything else.
 *
 * Multi-threaded with OpenMP over the batch; this is also the CPU baseline ("port").
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <omp.h>

#define QO_API __attribute__((visibility("default")))

typedef struct { uint32_t a, b; } qo_key;

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
static inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

/* ------------------------------------------------------------------ Threefry-2x32-20 */
static inline uint32_t rotl32(uint32_t x, int r) { return (x << r) | (x >> (32 - r)); }

static inline void threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1,
                                uint32_t* o0, uint32_t* o1) {
    static const int R[2][4] = {{13, 15, 26, 6}, {17, 29, 16, 24}};
    uint32_t ks[3] = {k0, k1, k0 ^ k1 ^ 0x1BD11BDAu};
    uint32_t x0 = c0 + ks[0], x1 = c1 + ks[1];
    for (int g = 0; g < 5; ++g) {
        for (int r = 0; r < 4; ++r) {
            x0 += x1;
            x1 = rotl32(x1, R[g & 1][r]);
            x1 ^= x0;
        }
        x0 += ks[(g + 1) % 3];
        x1 += ks[(g + 2) % 3] + (uint32_t)(g + 1);
    }
    *o0 = x0; *o1 = x1;
}

/* jax.random.split(key, n)[i] in partitionable mode */
static inline qo_key split_i(qo_key k, uint64_t i) {
    qo_key o;
    threefry2x32(k.a, k.b, (uint32_t)(i >> 32), (uint32_t)i, &o.a, &o.b);
    return o;
}
/* random_bits(key, 32, shape)[flat i] */
static inline uint32_t bits32(qo_key k, uint64_t i) {
    uint32_t a, b;
    threefry2x32(k.a, k.b, (uint32_t)(i >> 32), (uint32_t)i, &a, &b);
    return a ^ b;
}
static inline float unit_float(uint32_t bits) { return u2f((bits >> 9) | 0x3F800000u) - 1.0f; }
/* jnp.clip / jnp.maximum / jnp.minimum: NaN propagates (XLA max/min do); -0 < +0 (IEEE 754-2019 maximum/minimum). */
static inline float max_nanprop(float x, float lo) {
    if (x != x || lo != lo) return NAN;
    if (x == lo) return signbit(x) ? lo : x;
    return x < lo ? lo : x;
}
static inline float min_nanprop(float x, float hi) {
    if (x != x || hi != hi) return NAN;
    if (x == hi) return signbit(x) ? x : hi;
    return x > hi ? hi : x;
}

/* ------------------------------------------------------------------ QDX-F32 spec math */
/* log(t), t > 0 normal.  t = m * 2^e, m in [sqrt(1/2), sqrt(2)); Cephes-style degree-9 kernel. */
static inline float spec_logf(float t) {
    uint32_t ix = f2u(t) - 0x3f3504f3u;
    int e = (int32_t)ix >> 23;
    float m = u2f((ix & 0x007fffffu) + 0x3f3504f3u);
    float r = m - 1.0f;
    float z = r * r;
    float p = 7.0376836292E-2f;
    p = fmaf(p, r, -1.1514610310E-1f);
    p = fmaf(p, r, 1.1676998740E-1f);
    p = fmaf(p, r, -1.2420140846E-1f);
    p = fmaf(p, r, 1.4249322787E-1f);
    p = fmaf(p, r, -1.6668057665E-1f);
    p = fmaf(p, r, 2.0000714765E-1f);
    p = fmaf(p, r, -2.4999993993E-1f);
    p = fmaf(p, r, 3.3333331174E-1f);
    float y = (p * r) * z;
    float fe = (float)e;
    y = fmaf(fe, -2.12194440e-4f, y);
    y = fmaf(z, -0.5f, y);
    float res = r + y;
    res = fmaf(fe, 0.693359375f, res);
    return res;
}
/* log1p(y) for y in (-1, 0]:  log(t) + (y - (t - 1)) * (2 - t)  with t = fl(1 + y).
 * (y - (t - 1)) is the rounding error of t (exact); (2 - t) stands in for 1/t: the correction is at most half an
 * ulp of t, and wherever 2 - t is a poor reciprocal (t << 1) |log t| > 0.69 dwarfs it.  t == 1 gives exactly y. */
static inline float spec_log1pf(float y) {
    float t = 1.0f + y;
    float c = (y - (t - 1.0f)) * (2.0f - t);
    return spec_logf(t) + c;
}
/* XLA ErfInv32 (Giles) with fused Horner steps. */
static inline float spec_erfinvf(float x) {
    float w = -spec_log1pf(-(x * x));
    float p;
    if (w < 5.0f) {
        w = w - 2.5f;
        p = 2.81022636e-08f;
        p = fmaf(p, w, 3.43273939e-07f);
        p = fmaf(p, w, -3.5233877e-06f);
        p = fmaf(p, w, -4.39150654e-06f);
        p = fmaf(p, w, 0.00021858087f);
        p = fmaf(p, w, -0.00125372503f);
        p = fmaf(p, w, -0.00417768164f);
        p = fmaf(p, w, 0.246640727f);
        p = fmaf(p, w, 1.50140941f);
    } else {
        w = sqrtf(w) - 3.0f;
        p = -0.000200214257f;
        p = fmaf(p, w, 0.000100950558f);
        p = fmaf(p, w, 0.00134934322f);
        p = fmaf(p, w, -0.00367342844f);
        p = fmaf(p, w, 0.00573950773f);
        p = fmaf(p, w, -0.0076224613f);
        p = fmaf(p, w, 0.00943887047f);
        p = fmaf(p, w, 1.00167406f);
        p = fmaf(p, w, 2.83297682f);
    }
    if (fabsf(x) == 1.0f) return x * 3.40282347e+38f;
    return p * x;
}
/* exp(z): n = rint(z*log2e), r = z - n*ln2 (two-term, fused), degree-6 polynomial, scale by 2^n (two steps so that
 * results down to the subnormal range and up to FLT_MAX are produced by plain multiplications). */
static inline float spec_expf(float z) {
    if (z != z) return z;
    if (z > 88.75f) return INFINITY;
    if (z < -104.0f) return 0.0f;
    float n = rintf(z * 0x1.715476p+0f);
    float r = fmaf(n, -0x1.62e400p-1f, z);
    r = fmaf(n, -0x1.7f7d1cp-20f, r);
    float p = 0x1.6c16c2p-10f;                 /* 1/720 */
    p = fmaf(p, r, 0x1.111112p-7f);            /* 1/120 */
    p = fmaf(p, r, 0x1.555556p-5f);            /* 1/24  */
    p = fmaf(p, r, 0x1.555556p-3f);            /* 1/6   */
    p = fmaf(p, r, 0.5f);
    p = fmaf(p, r, 1.0f);
    p = fmaf(p, r, 1.0f);
    int ni = (int)n;
    int h = ni / 2;
    float s1 = u2f((uint32_t)(h + 127) << 23), s2 = u2f((uint32_t)(ni - h + 127) << 23);
    return (p * s1) * s2;
}
/* pow(x, y) for the polynomial mutation (x >= 0 expected): exp(y * log(x)); pow(0, y > 0) = 0; x < 0 -> NaN. */
static inline float spec_powf(float x, float y) {
    if (x != x || y != y) return NAN;
    if (x < 0.0f) return NAN;
    if (x == 0.0f) return y > 0.0f ? 0.0f : (y == 0.0f ? 1.0f : INFINITY);
    if (x == INFINITY) return y > 0.0f ? INFINITY : (y == 0.0f ? 1.0f : 0.0f);
    float lg;
    if (x < 0x1p-126f) lg = spec_logf(x * 0x1p+25f) - 0x1.154246p+4f;   /* subnormal: log(x 2^25) - 25 ln 2 */
    else lg = spec_logf(x);
    return spec_expf(y * lg);
}
/* jax.random.normal from one 32-bit draw: sqrt(2) * erfinv(max(lo, f*2 + lo)), lo = nextafter(-1, 0). */
static inline float normal_from_bits(uint32_t bits) {
    const float lo = -0x1.fffffep-1f;
    float f = unit_float(bits);
    float u = f * 2.0f + lo;
    u = u < lo ? lo : u;
    return 0x1.6a09e6p+0f * spec_erfinvf(u);
}
/* sin and cos: 3-term Cody-Waite reduction by pi/2 (fused), Cephes minimax kernels on [-pi/4, pi/4]. */
static inline void spec_sincosf(float th, float* s_out, float* c_out) {
    float q = rintf(th * 0x1.45f306p-1f);
    float r = fmaf(q, -0x1.921fb6p+0f, th);
    r = fmaf(q, 0x1.777a5cp-25f, r);
    r = fmaf(q, 0x1.ee59dap-50f, r);
    int n = (int)q & 3;
    float s = r * r;
    float ps = fmaf(fmaf(-1.9515295891E-4f, s, 8.3321608736E-3f), s, -1.6666654611E-1f);
    float sr = fmaf(r * s, ps, r);
    float pc = fmaf(fmaf(2.443315711809948E-5f, s, -1.388731625493765E-3f), s, 4.166664568298827E-2f);
    float cr = fmaf(s * s, pc, fmaf(s, -0.5f, 1.0f));
    float sv = (n & 1) ? cr : sr;
    float cv = (n & 1) ? sr : cr;
    if (n & 2) sv = -sv;
    if ((n + 1) & 2) cv = -cv;
    *s_out = sv; *c_out = cv;
}

/* ------------------------------------------------------------------ exported probes */
QO_API int qo_set_threads(int n) { if (n > 0) omp_set_num_threads(n); return 0; }
QO_API int qo_get_threads(void) { return omp_get_max_threads(); }

QO_API int qo_threefry2x32(uint32_t k0, uint32_t k1, uint32_t c0, uint32_t c1, uint32_t* out2) {
    threefry2x32(k0, k1, c0, c1, &out2[0], &out2[1]);
    return 0;
}
QO_API int qo_split(const uint32_t* key, int64_t num, uint32_t* out) {
    qo_key k = {key[0], key[1]};
    for (int64_t i = 0; i < num; ++i) { qo_key o = split_i(k, (uint64_t)i); out[2 * i] = o.a; out[2 * i + 1] = o.b; }
    return 0;
}
QO_API int qo_random_bits(const uint32_t* key, int64_t n, uint32_t* out) {
    qo_key k = {key[0], key[1]};
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = bits32(k, (uint64_t)i);
    return 0;
}
QO_API int qo_uniform(const uint32_t* key, int64_t n, float minval, float maxval, float* out) {
    qo_key k = {key[0], key[1]};
    float scale = maxval - minval;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        float v = unit_float(bits32(k, (uint64_t)i)) * scale + minval;
        out[i] = v < minval ? minval : v;
    }
    return 0;
}
QO_API int qo_normal(const uint32_t* key, int64_t n, float* out) {
    qo_key k = {key[0], key[1]};
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) out[i] = normal_from_bits(bits32(k, (uint64_t)i));
    return 0;
}
/* which: 0 log1p(x in (-1,0]), 1 erfinv, 2 sin, 3 cos */
QO_API int qo_math_probe(int which, int64_t n, const float* in, float* out) {
    for (int64_t i = 0; i < n; ++i) {
        float s, c;
        switch (which) {
            case 0: out[i] = spec_log1pf(in[i]); break;
            case 1: out[i] = spec_erfinvf(in[i]); break;
            case 2: spec_sincosf(in[i], &s, &c); out[i] = s; break;
            case 3: spec_sincosf(in[i], &s, &c); out[i] = c; break;
            default: return -1;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ selection
 * UniformSelector.select -- qdax/core/emitters/repertoire_selectors/uniform_selector.py:22-62.
 * `key` is the key handed to select(); :48 splits it once.  jax.random.choice(p):
 * cum = cumsum(p) (sequential float32), r = cum[-1] * (1 - u), searchsorted(cum, r, 'left'). */
static int select_indices(const float* fit, int64_t K, qo_key key, int64_t num, int32_t* out) {
    float* cum = (float*)malloc(sizeof(float) * (size_t)K);
    if (!cum) return -2;
    int64_t M = 0;
    for (int64_t c = 0; c < K; ++c) M += (fit[c] != -INFINITY);
    if (M == 0) { free(cum); return -3; }  /* p = 0/0 in the reference: undefined */
    float q = 1.0f / (float)M;
    float acc = 0.0f;
    for (int64_t c = 0; c < K; ++c) { acc = acc + ((fit[c] != -INFINITY) ? q : 0.0f); cum[c] = acc; }
    float total = cum[K - 1];
    qo_key sub = split_i(key, 1);
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < num; ++i) {
        float u = unit_float(bits32(sub, (uint64_t)i));
        float r = total * (1.0f - u);
        int64_t lo = 0, hi = K;                 /* first index with cum[idx] >= r */
        while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (cum[mid] < r) lo = mid + 1; else hi = mid; }
        out[i] = (int32_t)lo;
    }
    free(cum);
    return 0;
}
QO_API int qo_select_indices(const float* fit, int64_t K, const uint32_t* key, int64_t num, int32_t* out) {
    qo_key k = {key[0], key[1]};
    return select_indices(fit, K, k, num, out);
}

/* UniformSelector(select_with_replacement=False) -- jax.random.choice(..., p=p, replace=False): Gumbel top-k.
 * g_c = -log(-log(u_c)) + log(p_c), u = uniform(subkey, (K,), minval=tiny, maxval=1); indices = top_k(g, n), equal values:
 * lower index first; empty cells (p = 0, log = -inf) come last. */
typedef struct { float g; int32_t c; } qo_gc;
static int gc_cmp(const void* a, const void* b) {
    const qo_gc* x = (const qo_gc*)a; const qo_gc* y = (const qo_gc*)b;
    if (x->g > y->g) return -1;
    if (x->g < y->g) return 1;
    return (x->c > y->c) - (x->c < y->c);
}
QO_API int qo_select_indices_noreplace(const float* fit, int64_t K, const uint32_t* key, int64_t num, int32_t* out) {
    if (num > K) return -1;
    int64_t M = 0;
    for (int64_t c = 0; c < K; ++c) M += (fit[c] != -INFINITY);
    if (M == 0) return -3;
    qo_gc* v = (qo_gc*)malloc(sizeof(qo_gc) * (size_t)K);
    if (!v) return -2;
    qo_key k = {key[0], key[1]};
    qo_key sub = split_i(k, 1);
    const float q = 1.0f / (float)M;
    const float logq = spec_logf(q);
    const float tiny = 0x1p-126f;
    for (int64_t c = 0; c < K; ++c) {
        float f = unit_float(bits32(sub, (uint64_t)c));
        float u = f * 1.0f + tiny;                      /* f * (maxval - minval) + minval, maxval - minval = fl(1 - tiny) = 1 */
        if (u < tiny) u = tiny;
        float g = -spec_logf(-spec_logf(u));
        v[c].g = (fit[c] != -INFINITY) ? g + logq : -INFINITY;
        v[c].c = (int32_t)c;
    }
    qsort(v, (size_t)K, sizeof(qo_gc), gc_cmp);
    for (int64_t i = 0; i < num; ++i) out[i] = v[i].c;
    free(v);
    return 0;
}

/* ------------------------------------------------------------------ isoline variation
 * qdax/core/emitters/mutation_operators.py:175-226 (single-leaf genotype).
 * x = (x1 + iso) + (x2 - x1) * line ; clip.  Parents given by index into a (K, D) table when
 * p1/p2 are non-NULL, else x1/x2 are dense (B, D). */
/* Pytree genotypes (:219-224): an individual is the concatenation of its flattened leaves (jax.tree.leaves order), leaf l
 * owning genes [off[l], off[l+1]) of the packed row; keys = split(key, nb_leaves); leaf l's noise is
 * normal(keys[l], (B,) + leaf_shape), i.e. counter i * size_l + j.  n_leaves <= 1 / off == NULL: one leaf of D genes. */
static void isoline_leaves(const float* x1, const float* x2, const int32_t* p1, const int32_t* p2, int64_t B, int64_t D,
                           qo_key key, int n_leaves, const int32_t* off, float iso_sigma, float line_sigma, int has_min, float minv,
                           int has_max, float maxv, float* out) {
    qo_key k_line = split_i(key, 1);             /* :205  key, key_line_noise = split(key) */
    qo_key k_rest = split_i(key, 0);
    const int32_t one_off[2] = {0, (int32_t)D};
    if (n_leaves <= 1 || !off) { n_leaves = 1; off = one_off; }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < B; ++i) {
        float line = normal_from_bits(bits32(k_line, (uint64_t)i)) * line_sigma;   /* :207 */
        const float* a = p1 ? x1 + (int64_t)p1[i] * D : x1 + i * D;
        const float* b = p2 ? x2 + (int64_t)p2[i] * D : x2 + i * D;
        for (int l = 0; l < n_leaves; ++l) {
            const qo_key k_leaf = split_i(k_rest, (uint64_t)l);          /* :220  split(key, nb_leaves)[l] */
            const int64_t sz = off[l + 1] - off[l];
            for (int64_t j = 0; j < sz; ++j) {
                const int64_t d = off[l] + j;
                float iso = normal_from_bits(bits32(k_leaf, (uint64_t)(i * sz + j))) * iso_sigma;   /* :210 */
                float t1 = a[d] + iso;
                float t2 = b[d] - a[d];
                float t3 = t2 * line;
                float x = t1 + t3;                                                             /* :211 */
                if (has_min) x = max_nanprop(x, minv);                                              /* :214-215 */
                if (has_max) x = min_nanprop(x, maxv);
                out[i * D + d] = x;
            }
        }
    }
}
static void isoline(const float* x1, const float* x2, const int32_t* p1, const int32_t* p2, int64_t B, int64_t D,
                    qo_key key, float iso_sigma, float line_sigma, int has_min, float minv, int has_max, float maxv,
                    float* out) {
    qo_key k_line = split_i(key, 1);             /* :205  key, key_line_noise = split(key) */
    qo_key k_rest = split_i(key, 0);
    qo_key k_leaf = split_i(k_rest, 0);          /* :220  split(key, nb_leaves=1)[0] */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < B; ++i) {
        float line = normal_from_bits(bits32(k_line, (uint64_t)i)) * line_sigma;   /* :207 */
        const float* a = p1 ? x1 + (int64_t)p1[i] * D : x1 + i * D;
        const float* b = p2 ? x2 + (int64_t)p2[i] * D : x2 + i * D;
        for (int64_t d = 0; d < D; ++d) {
            float iso = normal_from_bits(bits32(k_leaf, (uint64_t)(i * D + d))) * iso_sigma;   /* :210 */
            float t1 = a[d] + iso;
            float t2 = b[d] - a[d];
            float t3 = t2 * line;
            float x = t1 + t3;                                                             /* :211 */
            if (has_min) x = max_nanprop(x, minv);                                              /* :214-215 */
            if (has_max) x = min_nanprop(x, maxv);
            out[i * D + d] = x;
        }
    }
}
QO_API int qo_isoline_variation(const float* x1, const float* x2, int64_t B, int64_t D, const uint32_t* key,
                                float iso_sigma, float line_sigma, int has_min, float minv, int has_max, float maxv,
                                float* out) {
    qo_key k = {key[0], key[1]};
    isoline(x1, x2, NULL, NULL, B, D, k, iso_sigma, line_sigma, has_min, minv, has_max, maxv, out);
    return 0;
}
QO_API int qo_isoline_variation_leaves(const float* x1, const float* x2, int64_t B, int64_t D, const uint32_t* key, int n_leaves,
                                       const int32_t* off, float iso_sigma, float line_sigma, int has_min, float minv, int has_max,
                                       float maxv, float* out) {
    qo_key k = {key[0], key[1]};
    if (n_leaves < 1 || !off || off[0] != 0 || off[n_leaves] != D) return -1;
    isoline_leaves(x1, x2, NULL, NULL, B, D, k, n_leaves, off, iso_sigma, line_sigma, has_min, minv, has_max, maxv, out);
    return 0;
}
/* ------------------------------------------------------------------ polynomial mutation / crossover
 * qdax/core/emitters/mutation_operators.py:12-117 and :120-172.  keys = split(key, B); per row:
 *   key, sub = split(key_r); positions = choice(sub, arange(D), (n,), replace=False) = permutation(sub, D)[:n]
 *   (permutation: `rounds` passes of  k, s = split(k); stable sort of the array by random_bits(s, (D,)) )
 *   key, sub = split(key); rand = uniform(sub, (n,)); polynomial delta; x[positions] += delta * (hi - lo); clip. */
static int perm_rounds(int64_t n) {
    if (n <= 1) return 1;
    return (int)ceil(3.0 * log((double)n) / log(4294967295.0));
}
typedef struct { uint32_t key; int32_t pos; int32_t val; } qo_pk;
static int pk_cmp(const void* a, const void* b) {
    const qo_pk* x = (const qo_pk*)a; const qo_pk* y = (const qo_pk*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);       /* stable */
}
static void permutation(qo_key key, int64_t n, int32_t* out, qo_pk* tmp) {
    for (int64_t i = 0; i < n; ++i) out[i] = (int32_t)i;
    int rounds = perm_rounds(n);
    for (int r = 0; r < rounds; ++r) {
        qo_key sub = split_i(key, 1); key = split_i(key, 0);
        for (int64_t i = 0; i < n; ++i) { tmp[i].key = bits32(sub, (uint64_t)i); tmp[i].pos = (int32_t)i; tmp[i].val = out[i]; }
        qsort(tmp, (size_t)n, sizeof(qo_pk), pk_cmp);
        for (int64_t i = 0; i < n; ++i) out[i] = tmp[i].val;
    }
}
QO_API int qo_polynomial_mutation(const float* x, int64_t B, int64_t D, const uint32_t* key, float proportion, float eta,
                                  float minv, float maxv, float* out) {
    qo_key k = {key[0], key[1]};
    const int64_t n = (int64_t)((double)proportion * (double)D);     /* int(proportion_to_mutate * num_positions)  :40 */
    const float rng = maxv - minv, mutpow = (float)(1.0 / (1.0 + (double)eta)), ep1 = (float)(1.0 + (double)eta);
    int rc = 0;
#pragma omp parallel
    {
        int32_t* perm = (int32_t*)malloc(sizeof(int32_t) * (size_t)D);
        qo_pk* tmp = (qo_pk*)malloc(sizeof(qo_pk) * (size_t)D);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < B; ++i) {
            if (!perm || !tmp) { rc = -2; continue; }
            qo_key kr = split_i(k, (uint64_t)i);                    /* :107 */
            qo_key sub1 = split_i(kr, 1), k1 = split_i(kr, 0);      /* :41 */
            permutation(sub1, D, perm, tmp);                        /* :42-45 */
            qo_key sub2 = split_i(k1, 1);                           /* :53 */
            float* o = out + i * D;
            for (int64_t d = 0; d < D; ++d) o[d] = x[i * D + d];
            for (int64_t j = 0; j < n; ++j) {
                const int32_t pos = perm[j];
                const float mx = x[i * D + pos];
                const float d1 = (mx - minv) / rng, d2 = (maxv - mx) / rng;
                const float r = unit_float(bits32(sub2, (uint64_t)j));              /* :54-60 */
                float v1 = 2.0f * r + spec_powf(d1, ep1) * (1.0f - 2.0f * r);         /* :62 */
                float v2 = 2.0f * (1.0f - r) + 2.0f * (spec_powf(d2, ep1) * (r - 0.5f));   /* :63 */
                v1 = spec_powf(v1, mutpow) - 1.0f;                                  /* :64 */
                v2 = 1.0f - spec_powf(v2, mutpow);                                  /* :65 */
                const float dq = r < 0.5f ? v1 : v2;                                /* :67-69 */
                o[pos] = mx + dq * rng;                                             /* :72 */
            }
            for (int64_t d = 0; d < D; ++d) o[d] = min_nanprop(max_nanprop(o[d], minv), maxv);   /* :75 */
        }
        free(perm); free(tmp);
    }
    return rc;
}
/* jax.random.randint(key, (n,), 0, span): two 32-bit draws per element from split(key), combined modulo span in uint32 */
static inline uint32_t randint_one(qo_key k1, qo_key k2, uint64_t j, uint32_t span) {
    uint32_t hi = bits32(k1, j), lo = bits32(k2, j);
    uint32_t mult = (65536u % span); mult = (mult * mult) % span;
    return ((hi % span) * mult + (lo % span)) % span;
}
QO_API int qo_polynomial_crossover(const float* x1, const float* x2, int64_t B, int64_t D, const uint32_t* key, float proportion,
                                   float* out) {
    qo_key k = {key[0], key[1]};
    const int64_t n = (int64_t)((double)proportion * (double)D);     /* :130 */
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < B; ++i) {
        qo_key kr = split_i(k, (uint64_t)i);                        /* :164 */
        qo_key r1 = split_i(kr, 0), r2 = split_i(kr, 1);            /* randint: k1, k2 = split(key) */
        for (int64_t d = 0; d < D; ++d) out[i * D + d] = x1[i * D + d];
        for (int64_t j = 0; j < n; ++j) {
            uint32_t idx = randint_one(r1, r2, (uint64_t)j, (uint32_t)D);           /* :132 */
            out[i * D + idx] = x2[i * D + idx];                                      /* :133 */
        }
    }
    return 0;
}
/* which: 4 exp, 5 pow(in, aux) */
QO_API int qo_math_probe2(int which, int64_t n, const float* in, float aux, float* out) {
    for (int64_t i = 0; i < n; ++i) {
        if (which == 4) out[i] = spec_expf(in[i]);
        else if (which == 5) out[i] = spec_powf(in[i], aux);
        else return -1;
    }
    return 0;
}

/* MixingEmitter.emit with variation_percentage = 1 -- standard_emitters.py:51-62 */
QO_API int qo_emit_isoline(const float* rep_g, const float* rep_f, int64_t K, int64_t D, const uint32_t* key, int64_t B,
                           float iso_sigma, float line_sigma, int has_min, float minv, int has_max, float maxv,
                           float* out, int32_t* p1, int32_t* p2) {
    qo_key k = {key[0], key[1]};
    int rc = select_indices(rep_f, K, split_i(k, 0), B, p1);  /* :55-56 */
    if (rc) return rc;
    rc = select_indices(rep_f, K, split_i(k, 1), B, p2);      /* :59 */
    if (rc) return rc;
    isoline(rep_g, rep_g, p1, p2, B, D, split_i(k, 2), iso_sigma, line_sigma, has_min, minv, has_max, maxv, out);
    return 0;
}

/* the same for a pytree genotype stored as packed rows */
QO_API int qo_emit_isoline_leaves(const float* rep_g, const float* rep_f, int64_t K, int64_t D, const uint32_t* key, int64_t B,
                                  int n_leaves, const int32_t* off, float iso_sigma, float line_sigma, int has_min, float minv,
                                  int has_max, float maxv, float* out, int32_t* p1, int32_t* p2) {
    qo_key k = {key[0], key[1]};
    if (n_leaves < 1 || !off || off[0] != 0 || off[n_leaves] != D) return -1;
    int rc = select_indices(rep_f, K, split_i(k, 0), B, p1);
    if (rc) return rc;
    rc = select_indices(rep_f, K, split_i(k, 1), B, p2);
    if (rc) return rc;
    isoline_leaves(rep_g, rep_g, p1, p2, B, D, split_i(k, 2), n_leaves, off, iso_sigma, line_sigma, has_min, minv, has_max, maxv, out);
    return 0;
}

/* ------------------------------------------------------------------ MELS reduction
 * qdax/core/containers/mels_repertoire.py: _mode :51-57 (smallest most frequent cell), _dispersion :26-48 (mean pairwise
 * distance, 0 when S == 1 per :152-158), mean fitness :169.  Sums sequential, left to right. */
QO_API int qo_mels_reduce(const int32_t* cells_all, const float* desc_all, const float* fit_all, int64_t B, int64_t S, int64_t Dd,
                          int32_t* out_cell, float* out_spread, float* out_fmean) {
    if (S < 1 || Dd < 1) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        const int32_t* c = cells_all + b * S;
        int32_t mode = c[0]; int64_t best = 0;
        for (int64_t i = 0; i < S; ++i) {
            int64_t n = 0;
            for (int64_t j = 0; j < S; ++j) n += (c[j] == c[i]);
            if (n > best || (n == best && c[i] < mode)) { best = n; mode = c[i]; }
        }
        const float* d = desc_all + b * S * Dd;
        float spread = 0.0f;
        if (S > 1) {
            float sum = 0.0f;
            for (int64_t i = 0; i < S; ++i)
                for (int64_t j = i + 1; j < S; ++j) {
                    float acc = 0.0f;
                    for (int64_t k = 0; k < Dd; ++k) { float t = d[i * Dd + k] - d[j * Dd + k]; acc = acc + t * t; }
                    sum = sum + sqrtf(acc);
                }
            spread = sum / (float)((double)S * (double)(S - 1) / 2.0);
        }
        float fs = 0.0f;
        for (int64_t i = 0; i < S; ++i) fs = fs + fit_all[b * S + i];
        out_cell[b] = mode; out_spread[b] = spread; out_fmean[b] = fs / (float)S;
    }
    return 0;
}

/* ------------------------------------------------------------------ scoring
 * task 0: arm (qdax/tasks/arm.py:9-38), 1: rastrigin, 2: sphere (qdax/tasks/standard_functions.py:9-24).
 * Descriptor of rastrigin/sphere = first Dd genes (Dd = 2 in the reference; other Dd is the declared
 * C4 extension).  */
static void score_row(int task, const float* p, int64_t D, int64_t Dd, float* f_out, float* desc) {
    if (task == 0) {
        float sum = 0.0f;
        for (int64_t d = 0; d < D; ++d) { float x = min_nanprop(max_nanprop(p[d], 0.0f), 1.0f); sum = sum + x; }
        float mean = sum / (float)D;
        float sq = 0.0f, th = 0.0f, cs = 0.0f, sn = 0.0f;
        for (int64_t d = 0; d < D; ++d) {
            float x = min_nanprop(max_nanprop(p[d], 0.0f), 1.0f);
            float dev = x - mean;
            float dd = dev * dev;
            sq = sq + dd;
            float ang = 0x1.921fb6p+2f * x - 0x1.921fb6p+1f;     /* 2*pi*x - pi */
            th = th + ang;                                        /* cumsum (running sum from 0) */
            float s, c;
            spec_sincosf(th, &s, &c);
            cs = cs + c;
            sn = sn + s;
        }
        float var = sq / (float)D;
        *f_out = -sqrtf(var);
        desc[0] = cs / (float)(2 * D) + 0.5f;
        desc[1] = sn / (float)(2 * D) + 0.5f;
    } else {
        float acc = 0.0f;
        for (int64_t d = 0; d < D; ++d) {
            float x = p[d] * 10.0f - 5.0f;
            float term = x * x;
            if (task == 1) {
                float s, c;
                spec_sincosf(0x1.921fb6p+2f * x, &s, &c);
                term = term - 10.0f * c;
            }
            acc = acc + term;
        }
        if (task == 1) acc = (float)(10.0 * (double)D) + acc;
        *f_out = -acc;
        for (int64_t j = 0; j < Dd; ++j) desc[j] = p[j];
    }
}
QO_API int qo_score(int task, const float* g, int64_t B, int64_t D, int64_t Dd, float* f, float* desc) {
    if (task < 0 || task > 2 || (task == 0 && Dd != 2) || Dd > D) return -1;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < B; ++i) score_row(task, g + i * D, D, Dd, f + i, desc + i * Dd);
    return 0;
}

/* noisy_arm_scoring_function -- qdax/tasks/arm.py:53-81: key, f_sub, d_sub, p_sub = split(key, 4); the genotype gets
 * normal(p_sub, params.shape) * params_variance added before the arm is evaluated, the fitness normal(f_sub, (B,)) *
 * fit_variance and the descriptor normal(d_sub, (B, 2)) * desc_variance afterwards (one rounding per operation). */
QO_API int qo_noisy_arm(const float* g, int64_t B, int64_t D, const uint32_t* key, float fit_variance, float desc_variance,
                        float params_variance, float* f, float* desc) {
    qo_key k = {key[0], key[1]};
    const qo_key kf = split_i(k, 1), kd = split_i(k, 2), kp = split_i(k, 3);
#pragma omp parallel
    {
        float* row = (float*)malloc(sizeof(float) * (size_t)D);
#pragma omp for schedule(static)
        for (int64_t i = 0; i < B; ++i) {
            for (int64_t d = 0; d < D; ++d) {
                float n = normal_from_bits(bits32(kp, (uint64_t)(i * D + d)));
                float t = n * params_variance;
                row[d] = g[i * D + d] + t;
            }
            float fi, di[2];
            score_row(0, row, D, 2, &fi, di);
            float nf = normal_from_bits(bits32(kf, (uint64_t)i)) * fit_variance;
            f[i] = fi + nf;
            for (int j = 0; j < 2; ++j) {
                float nd = normal_from_bits(bits32(kd, (uint64_t)(2 * i + j))) * desc_variance;
                desc[2 * i + j] = di[j] + nd;
            }
        }
        free(row);
    }
    return 0;
}

/* ------------------------------------------------------------------ cell assignment
 * get_cells_indices -- qdax/core/containers/mapelites_repertoire.py:111-137: brute force,
 * argmin_k sum_d (x_d - c_kd)^2 (sequential over d), first minimum, NaN counts as minimal. */
QO_API int qo_cells(const float* desc, int64_t B, int64_t Dd, const float* cent, int64_t K, int32_t* cells) {
    float* ct = (float*)malloc(sizeof(float) * (size_t)(K * Dd));   /* (Dd, K) transpose */
    if (!ct) return -2;
    for (int64_t k = 0; k < K; ++k) for (int64_t j = 0; j < Dd; ++j) ct[j * K + k] = cent[k * Dd + j];
#pragma omp parallel
    {
        enum { BLK = 1024 };
        float acc[BLK];
#pragma omp for schedule(static)
        for (int64_t i = 0; i < B; ++i) {
            const float* x = desc + i * Dd;
            float best = INFINITY; int64_t bk = -1; int nan_hit = 0;
            for (int64_t k0 = 0; k0 < K && !nan_hit; k0 += BLK) {
                int64_t n = K - k0 < BLK ? K - k0 : BLK;
                for (int64_t j = 0; j < Dd; ++j) {
                    const float xj = x[j]; const float* c = ct + j * K + k0;
                    if (j == 0) for (int64_t t = 0; t < n; ++t) { float df = xj - c[t]; acc[t] = df * df; }
                    else for (int64_t t = 0; t < n; ++t) { float df = xj - c[t]; acc[t] = acc[t] + df * df; }
                }
                for (int64_t t = 0; t < n; ++t) {
                    float v = acc[t];
                    if (v != v) { bk = k0 + t; nan_hit = 1; break; }     /* first NaN wins */
                    if (bk < 0 || v < best) { best = v; bk = k0 + t; }
                }
            }
            cells[i] = (int32_t)bk;
        }
    }
    free(ct);
    return 0;
}

/* ------------------------------------------------------------------ insertion
 * MapElitesRepertoire.add -- mapelites_repertoire.py:173-266, given precomputed cells.
 * tie_break_last = 0: among equal-fitness winners of one cell the FIRST offspring index is stored. */
QO_API int qo_add(float* rep_g, float* rep_f, float* rep_d, int64_t K, int64_t D, int64_t Dd, const float* g,
                  const float* f, const float* desc, const int32_t* cells, int64_t B, int tie_break_last,
                  int32_t* scatter_idx) {
    float* best = (float*)malloc(sizeof(float) * (size_t)K);
    int32_t* win = (int32_t*)malloc(sizeof(int32_t) * (size_t)K);
    if (!best || !win) { free(best); free(win); return -2; }
    for (int64_t c = 0; c < K; ++c) { best[c] = -INFINITY; win[c] = -1; }
    for (int64_t i = 0; i < B; ++i) {                       /* segment_max, NaN-propagating  :211-215 */
        int32_t c = cells[i];
        if (c < 0 || c >= K) { free(best); free(win); return -1; }
        float v = f[i], b = best[c];
        if (b != b) continue;
        if (v != v || v > b) best[c] = v;
    }
    for (int64_t i = 0; i < B; ++i) {
        int32_t c = cells[i];
        float fm = (f[i] == best[c]) ? f[i] : -INFINITY;    /* :217-222 */
        int cond = fm > rep_f[c];                            /* :225-226 strict */
        if (scatter_idx) scatter_idx[i] = cond ? c : (int32_t)K;   /* :229-231 */
        if (cond && (tie_break_last || win[c] < 0)) win[c] = (int32_t)i;
    }
    for (int64_t c = 0; c < K; ++c) {                        /* .at[idx].set  :234-257 */
        int32_t i = win[c];
        if (i < 0) continue;
        memcpy(rep_g + c * D, g + (int64_t)i * D, sizeof(float) * (size_t)D);
        rep_f[c] = f[i];
        memcpy(rep_d + c * Dd, desc + (int64_t)i * Dd, sizeof(float) * (size_t)Dd);
    }
    free(best); free(win);
    return 0;
}

/* default_qd_metrics -- qdax/utils/metrics.py:74-98.  out = {qd_score, max_fitness, coverage}. */
QO_API int qo_metrics(const float* rep_f, int64_t K, float qd_offset, float* out3) {
    double s = 0.0; int64_t filled = 0; float mx = -INFINITY; int nan = 0;
    for (int64_t c = 0; c < K; ++c) {
        float v = rep_f[c];
        if (v != -INFINITY) { s += (double)v; ++filled; }
        if (v != v) nan = 1; else if (v > mx) mx = v;
    }
    out3[0] = (float)s + qd_offset * (float)filled;
    out3[1] = nan ? NAN : mx;
    out3[2] = 100.0f * ((float)filled / (float)K);
    return 0;
}

/* ------------------------------------------------------------------ driver
 * lax.scan(MAPElites.scan_update) -- qdax/core/map_elites.py:148-225 with MixingEmitter(pct=1) + isoline.
 * key chain: scan_update :214 -> update :177 -> ask :241 -> emit ; :181 scoring key (unused by the tasks).
 * stage_seconds (optional, 4 doubles): emit, score, cells, add+metrics wall-clock. */
QO_API int qo_map_elites_scan(float* rep_g, float* rep_f, float* rep_d, const float* cent, int64_t K, int64_t D,
                              int64_t Dd, uint32_t* key_io, int64_t n_iter, int64_t B, int task, float iso_sigma,
                              float line_sigma, int has_min, float minv, int has_max, float maxv, int tie_break_last,
                              float qd_offset, float* metrics_out, double* stage_seconds) {
    float* x = (float*)malloc(sizeof(float) * (size_t)(B * D));
    float* f = (float*)malloc(sizeof(float) * (size_t)B);
    float* ds = (float*)malloc(sizeof(float) * (size_t)(B * Dd));
    int32_t* p1 = (int32_t*)malloc(sizeof(int32_t) * (size_t)B);
    int32_t* p2 = (int32_t*)malloc(sizeof(int32_t) * (size_t)B);
    int32_t* cells = (int32_t*)malloc(sizeof(int32_t) * (size_t)B);
    int rc = (x && f && ds && p1 && p2 && cells) ? 0 : -2;
    qo_key key = {key_io[0], key_io[1]};
    double t[4] = {0, 0, 0, 0};
    for (int64_t it = 0; it < n_iter && rc == 0; ++it) {
        qo_key sub = split_i(key, 1); key = split_i(key, 0);     /* :214 */
        qo_key ask_key = split_i(sub, 1);                         /* :177 */
        qo_key emit_key = split_i(ask_key, 1);                    /* :241 */
        uint32_t ek[2] = {emit_key.a, emit_key.b};
        double t0 = omp_get_wtime();
        rc = qo_emit_isoline(rep_g, rep_f, K, D, ek, B, iso_sigma, line_sigma, has_min, minv, has_max, maxv, x, p1, p2);
        if (rc) break;
        double t1 = omp_get_wtime();
        rc = qo_score(task, x, B, D, Dd, f, ds);
        if (rc) break;
        double t2 = omp_get_wtime();
        rc = qo_cells(ds, B, Dd, cent, K, cells);
        if (rc) break;
        double t3 = omp_get_wtime();
        rc = qo_add(rep_g, rep_f, rep_d, K, D, Dd, x, f, ds, cells, B, tie_break_last, NULL);
        if (rc) break;
        if (metrics_out) qo_metrics(rep_f, K, qd_offset, metrics_out + 3 * it);
        double t4 = omp_get_wtime();
        t[0] += t1 - t0; t[1] += t2 - t1; t[2] += t3 - t2; t[3] += t4 - t3;
    }
    key_io[0] = key.a; key_io[1] = key.b;
    if (stage_seconds) memcpy(stage_seconds, t, sizeof(t));
    free(x); free(f); free(ds); free(p1); free(p2); free(cells);
    return rc;
}

/* DistributedMAPElites.update -- qdax/core/distributed_map_elites.py:92-161 for R simulated devices.
 * keys: R*2 words (the per-device key handed to update).  B_dev offspring per device; global offspring
 * index = rank * B_dev + i (all_gather + concatenate(axis=0), :134-141). */
QO_API int qo_distributed_update(float* rep_g, float* rep_f, float* rep_d, const float* cent, int64_t K, int64_t D,
                                 int64_t Dd, const uint32_t* keys, int64_t R, int64_t B_dev, int task,
                                 float iso_sigma, float line_sigma, int has_min, float minv, int has_max, float maxv,
                                 int tie_break_last, float* g_out, float* f_out, float* d_out, int32_t* cells_out) {
    int64_t B = R * B_dev;
    int32_t* p1 = (int32_t*)malloc(sizeof(int32_t) * (size_t)B_dev);
    int32_t* p2 = (int32_t*)malloc(sizeof(int32_t) * (size_t)B_dev);
    int rc = (p1 && p2) ? 0 : -2;
    for (int64_t r = 0; r < R && rc == 0; ++r) {
        qo_key key = {keys[2 * r], keys[2 * r + 1]};
        qo_key emit_key = split_i(key, 1);                        /* :124 */
        uint32_t ek[2] = {emit_key.a, emit_key.b};
        rc = qo_emit_isoline(rep_g, rep_f, K, D, ek, B_dev, iso_sigma, line_sigma, has_min, minv, has_max, maxv,
                             g_out + r * B_dev * D, p1, p2);
        if (rc) break;
        rc = qo_score(task, g_out + r * B_dev * D, B_dev, D, Dd, f_out + r * B_dev, d_out + r * B_dev * Dd);
    }
    if (rc == 0) rc = qo_cells(d_out, B, Dd, cent, K, cells_out);
    if (rc == 0) rc = qo_add(rep_g, rep_f, rep_d, K, D, Dd, g_out, f_out, d_out, cells_out, B, tie_break_last, NULL);
    free(p1); free(p2);
    return rc;
}

/* ------------------------------------------------------------------ Dominated Novelty Search
 * _novelty_and_dominated_novelty -- qdax/core/containers/dns_repertoire.py:22-76 (dominated novelty
 * only; add discards plain novelty, :136).  dn_i = mean of the k smallest sqrt(sum_d (x_id - x_jd)^2)
 * over j != i with both valid and f_i <= f_j; fewer than k such j -> mean over those; none -> NaN. */
#define QO_KMAX 32
QO_API int qo_dns_dominated_novelty(const float* f, const float* desc, int64_t N, int64_t Dd, int k, float* out) {
    if (k < 1 || k > QO_KMAX) return -1;
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < N; ++i) {
        float top[QO_KMAX]; int cnt = 0;
        float fi = f[i];
        if (fi == -INFINITY) { out[i] = NAN; continue; }
        for (int64_t j = 0; j < N; ++j) {
            if (j == i || f[j] == -INFINITY || !(fi <= f[j])) continue;
            float acc = 0.0f;
            for (int64_t d = 0; d < Dd; ++d) { float df = desc[i * Dd + d] - desc[j * Dd + d]; float s = df * df; acc = d ? acc + s : s; }
            float dist = sqrtf(acc);
            /* keep k smallest, ascending; NaN/inf distances sort last */
            if (cnt < k) { int p = cnt++; while (p > 0 && dist < top[p - 1]) { top[p] = top[p - 1]; --p; } top[p] = dist; }
            else if (dist < top[k - 1]) { int p = k - 1; while (p > 0 && dist < top[p - 1]) { top[p] = top[p - 1]; --p; } top[p] = dist; }
        }
        float tot = 0.0f;
        for (int j = 0; j < cnt; ++j) tot = tot + top[j];
        out[i] = tot / (float)cnt;          /* 0/0 -> NaN */
    }
    return 0;
}
/* survivor order = argsort(meta)[::-1]: NaN first, then descending, higher index first among equals (:148). */
typedef struct { uint32_t key; int32_t idx; } qo_sk;
static int sk_cmp(const void* a, const void* b) {
    const qo_sk* x = (const qo_sk*)a; const qo_sk* y = (const qo_sk*)b;
    if (x->key != y->key) return x->key > y->key ? -1 : 1;
    return x->idx > y->idx ? -1 : (x->idx < y->idx ? 1 : 0);
}
static inline uint32_t order_key(float v) {     /* total order: -inf < ... < -0 == +0 < ... < +inf < NaN */
    if (v != v) return 0xFFFFFFFFu;
    uint32_t u = f2u(v == 0.0f ? 0.0f : v);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
QO_API int qo_dns_survivors(const float* meta, int64_t N, int64_t P, int32_t* out) {
    qo_sk* a = (qo_sk*)malloc(sizeof(qo_sk) * (size_t)N);
    if (!a) return -2;
    for (int64_t i = 0; i < N; ++i) { a[i].key = order_key(meta[i]); a[i].idx = (int32_t)i; }
    qsort(a, (size_t)N, sizeof(qo_sk), sk_cmp);
    for (int64_t i = 0; i < P && i < N; ++i) out[i] = a[i].idx;
    free(a);
    return 0;
}
