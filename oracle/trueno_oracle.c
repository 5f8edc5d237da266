/*
 * trueno_oracle.c — CPU restatement of paiml/trueno's hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * This file is the parity oracle for the B200 kernels.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load it; the product (trueno_b200/,
 * include/) never links, imports or falls back to it.
 *
 * The reference is Rust and cannot be built in this image (no cargo/rustc), so each function
 * below restates the arithmetic AND the accumulation order of the reference function it cites
 * (paths relative to /root/reference).  All hot-path arithmetic is in-tree in the reference;
 * the only external piece is the platform libm (Rust's f32::exp/ln/tanh/sqrt lower to glibc
 * expf/logf/tanhf/sqrtf on linux-gnu — the same functions called here).
 *
 * Parity pin: tests/test_oracle_kat.py checks every function against the known-answer tests
 * and seeded fixtures the reference's own test-suite holds for this path.
 *
 * Build: see oracle/Makefile.  -ffp-contract=off is REQUIRED: Rust never contracts a*b+c into
 * an FMA, so every fused operation here is spelled fmaf()/_mm256_fmadd_ps explicitly, exactly
 * where the reference spells mul_add/_mm256_fmadd_ps.
 */
#include <immintrin.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* status codes == include/trueno_cuda.h trn_status (TruenoError variants, src/error.rs:8-41) */
enum { ORC_OK = 0, ORC_SIZE_MISMATCH = 1, ORC_INVALID_INPUT = 2, ORC_EMPTY_VECTOR = 3 };

static __thread char g_msg[512];
static __thread uint64_t g_expected, g_actual;

ORC_API const char* orc_last_message(void) { return g_msg; }
ORC_API void orc_last_mismatch(uint64_t* e, uint64_t* a) { *e = g_expected; *a = g_actual; }

static int fail_invalid(const char* m) {
    snprintf(g_msg, sizeof g_msg, "%s", m);
    return ORC_INVALID_INPUT;
}
static int fail_size(size_t expected, size_t actual) {
    g_expected = expected; g_actual = actual;
    snprintf(g_msg, sizeof g_msg, "Size mismatch: expected %zu, got %zu", expected, actual);
    return ORC_SIZE_MISMATCH;
}

/* ------------------------------------------------------------------------------------------
 * Scalar backend — the semantic ground truth (src/backends/scalar.rs)
 * ---------------------------------------------------------------------------------------- */

/* scalar.rs:66-94 — four fused-multiply-add chains, (c0+c1)+(c2+c3), fused tail */
ORC_API float orc_scalar_dot(const float* a, const float* b, size_t n) {
    float c0 = 0.f, c1 = 0.f, c2 = 0.f, c3 = 0.f;
    size_t q = n / 4;
    for (size_t i = 0; i < q; ++i) {
        const float* pa = a + 4 * i; const float* pb = b + 4 * i;
        c0 = fmaf(pa[0], pb[0], c0);
        c1 = fmaf(pa[1], pb[1], c1);
        c2 = fmaf(pa[2], pb[2], c2);
        c3 = fmaf(pa[3], pb[3], c3);
    }
    float s = (c0 + c1) + (c2 + c3);
    for (size_t i = 4 * q; i < n; ++i) s = fmaf(a[i], b[i], s);
    return s;
}

/* scalar.rs:100-106 — plain left-to-right f32 sum */
ORC_API float orc_scalar_sum(const float* a, size_t n) {
    float t = 0.f;
    for (size_t i = 0; i < n; ++i) t += a[i];
    return t;
}

/* scalar.rs:112-120 / :126-134 — seed with a[0], strict compare (a NaN never wins, a NaN seed never loses) */
ORC_API float orc_scalar_max(const float* a, size_t n) {
    float m = a[0];
    for (size_t i = 1; i < n; ++i) if (a[i] > m) m = a[i];
    return m;
}
ORC_API float orc_scalar_min(const float* a, size_t n) {
    float m = a[0];
    for (size_t i = 1; i < n; ++i) if (a[i] < m) m = a[i];
    return m;
}

/* scalar.rs:140-150 / :156-166 — first occurrence, strict compare, seed index 0 */
ORC_API uint64_t orc_scalar_argmax(const float* a, size_t n) {
    float m = a[0]; uint64_t at = 0;
    for (size_t i = 0; i < n; ++i) if (a[i] > m) { m = a[i]; at = i; }
    return at;
}
ORC_API uint64_t orc_scalar_argmin(const float* a, size_t n) {
    float m = a[0]; uint64_t at = 0;
    for (size_t i = 0; i < n; ++i) if (a[i] < m) { m = a[i]; at = i; }
    return at;
}

/* scalar.rs:190-200 — sequential sum of val*val (mul, then add), sqrt */
ORC_API float orc_scalar_norm_l2(const float* a, size_t n) {
    if (n == 0) return 0.f;
    float ss = 0.f;
    for (size_t i = 0; i < n; ++i) { float sq = a[i] * a[i]; ss += sq; }
    return sqrtf(ss);
}

static inline float sigmoid_libm(float v) {
    /* scalar.rs:313-324 — hard cut-offs at +-50, libm expf in between */
    if (v < -50.f) return 0.f;
    if (v > 50.f) return 1.f;
    return 1.f / (1.f + expf(-v));
}
static inline float gelu_libm(float x) {
    /* scalar.rs:330-340 — tanh approximation with libm tanhf; (x*x)*x, c*x3, x+.., k*(..) */
    const float k = 0.7978846f, c = 0.044715f;
    float x3 = x * x * x;
    float t = c * x3;
    float inner = k * (x + t);
    float h = 0.5f * x;
    return h * (1.f + tanhf(inner));
}
ORC_API void orc_scalar_sigmoid(const float* a, float* out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = sigmoid_libm(a[i]);
}
ORC_API void orc_scalar_gelu(const float* a, float* out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = gelu_libm(a[i]);
}
ORC_API void orc_scalar_add(const float* a, const float* b, float* out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = a[i] + b[i];
}
ORC_API void orc_scalar_mul(const float* a, const float* b, float* out, size_t n) {
    for (size_t i = 0; i < n; ++i) out[i] = a[i] * b[i];
}

/* ------------------------------------------------------------------------------------------
 * AVX2 backend — what trueno returns by default on x86 (src/backends/avx2.rs)
 * ---------------------------------------------------------------------------------------- */

/* lane fold used by dot and sum: lo128+hi128, then +movehl, then +lane1 (avx2.rs:203-210, :239-246) */
static inline float fold8_add(__m256 v) {
    __m128 h = _mm_add_ps(_mm256_castps256_ps128(v), _mm256_extractf128_ps(v, 1));
    __m128 t = _mm_add_ps(h, _mm_movehl_ps(h, h));
    t = _mm_add_ss(t, _mm_shuffle_ps(t, t, 1));
    return _mm_cvtss_f32(t);
}

/* avx2.rs:32-55 / :96-117 — one IEEE op per element: bit-exact with scalar */
ORC_API void orc_avx2_add(const float* a, const float* b, float* out, size_t n) {
    size_t i = 0;
    for (; i + 8 <= n; i += 8)
        _mm256_storeu_ps(out + i, _mm256_add_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i)));
    for (; i < n; ++i) out[i] = a[i] + b[i];
}
ORC_API void orc_avx2_mul(const float* a, const float* b, float* out, size_t n) {
    size_t i = 0;
    for (; i + 8 <= n; i += 8)
        _mm256_storeu_ps(out + i, _mm256_mul_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i)));
    for (; i < n; ++i) out[i] = a[i] * b[i];
}

/* avx2.rs:159-216 — 4 vector FMA chains over 32-element steps, leftover 8-steps go to chain 0,
 * (v0+v1)+(v2+v3), lane fold, then the scalar tail is summed SEPARATELY (mul, add from 0) and
 * added last. */
ORC_API float orc_avx2_dot(const float* a, const float* b, size_t n) {
    __m256 v0 = _mm256_setzero_ps(), v1 = v0, v2 = v0, v3 = v0;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        v0 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i),      _mm256_loadu_ps(b + i),      v0);
        v1 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 8),  _mm256_loadu_ps(b + i + 8),  v1);
        v2 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 16), _mm256_loadu_ps(b + i + 16), v2);
        v3 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i + 24), _mm256_loadu_ps(b + i + 24), v3);
    }
    for (; i + 8 <= n; i += 8)
        v0 = _mm256_fmadd_ps(_mm256_loadu_ps(a + i), _mm256_loadu_ps(b + i), v0);
    float r = fold8_add(_mm256_add_ps(_mm256_add_ps(v0, v1), _mm256_add_ps(v2, v3)));
    float tail = 0.f;
    for (; i < n; ++i) { float p = a[i] * b[i]; tail += p; }
    return r + tail;
}

/* avx2.rs:225-252 — ONE 8-lane accumulator (stagnates at large n: SURVEY.md top, fact 3) */
ORC_API float orc_avx2_sum(const float* a, size_t n) {
    __m256 acc = _mm256_setzero_ps();
    size_t i = 0;
    for (; i + 8 <= n; i += 8) acc = _mm256_add_ps(acc, _mm256_loadu_ps(a + i));
    float r = fold8_add(acc);
    float tail = 0.f;
    for (; i < n; ++i) tail += a[i];
    return r + tail;
}

/* avx2.rs:261-294 / :303-336 — lanes seeded with a[0]; vmaxps(acc, x) (returns x on NaN) */
ORC_API float orc_avx2_max(const float* a, size_t n) {
    __m256 m = _mm256_set1_ps(a[0]);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) m = _mm256_max_ps(m, _mm256_loadu_ps(a + i));
    __m128 h = _mm_max_ps(_mm256_castps256_ps128(m), _mm256_extractf128_ps(m, 1));
    __m128 t = _mm_max_ps(h, _mm_movehl_ps(h, h));
    t = _mm_max_ss(t, _mm_shuffle_ps(t, t, 1));
    float r = _mm_cvtss_f32(t);
    for (; i < n; ++i) if (a[i] > r) r = a[i];
    return r;
}
ORC_API float orc_avx2_min(const float* a, size_t n) {
    __m256 m = _mm256_set1_ps(a[0]);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) m = _mm256_min_ps(m, _mm256_loadu_ps(a + i));
    __m128 h = _mm_min_ps(_mm256_castps256_ps128(m), _mm256_extractf128_ps(m, 1));
    __m128 t = _mm_min_ps(h, _mm_movehl_ps(h, h));
    t = _mm_min_ss(t, _mm_shuffle_ps(t, t, 1));
    float r = _mm_cvtss_f32(t);
    for (; i < n; ++i) if (a[i] < r) r = a[i];
    return r;
}

/* avx2.rs:345-400 / :409-464 — indices carried as f32 lanes (lossy above 2^24 — a reference
 * DEFECT that the CUDA path does not reproduce; kept here to document what trueno returns). */
ORC_API uint64_t orc_avx2_argmax(const float* a, size_t n) {
    float best = a[0]; uint64_t at = 0;
    __m256 vb = _mm256_set1_ps(a[0]), vi = _mm256_setzero_ps();
    __m256 cur = _mm256_set_ps(7.f, 6.f, 5.f, 4.f, 3.f, 2.f, 1.f, 0.f);
    const __m256 step = _mm256_set1_ps(8.f);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        __m256 x = _mm256_loadu_ps(a + i);
        __m256 gt = _mm256_cmp_ps(x, vb, _CMP_GT_OQ);
        vb = _mm256_blendv_ps(vb, x, gt);
        vi = _mm256_blendv_ps(vi, cur, gt);
        cur = _mm256_add_ps(cur, step);
    }
    float lv[8], li[8];
    _mm256_storeu_ps(lv, vb); _mm256_storeu_ps(li, vi);
    for (int l = 0; l < 8; ++l) if (lv[l] > best) { best = lv[l]; at = (uint64_t)li[l]; }
    for (size_t j = i; j < n; ++j) if (a[j] > best) { best = a[j]; at = j; }
    return at;
}
ORC_API uint64_t orc_avx2_argmin(const float* a, size_t n) {
    float best = a[0]; uint64_t at = 0;
    __m256 vb = _mm256_set1_ps(a[0]), vi = _mm256_setzero_ps();
    __m256 cur = _mm256_set_ps(7.f, 6.f, 5.f, 4.f, 3.f, 2.f, 1.f, 0.f);
    const __m256 step = _mm256_set1_ps(8.f);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        __m256 x = _mm256_loadu_ps(a + i);
        __m256 lt = _mm256_cmp_ps(x, vb, _CMP_LT_OQ);
        vb = _mm256_blendv_ps(vb, x, lt);
        vi = _mm256_blendv_ps(vi, cur, lt);
        cur = _mm256_add_ps(cur, step);
    }
    float lv[8], li[8];
    _mm256_storeu_ps(lv, vb); _mm256_storeu_ps(li, vi);
    for (int l = 0; l < 8; ++l) if (lv[l] < best) { best = lv[l]; at = (uint64_t)li[l]; }
    for (size_t j = i; j < n; ++j) if (a[j] < best) { best = a[j]; at = j; }
    return at;
}

/* avx2.rs:481-489 */
ORC_API float orc_avx2_norm_l2(const float* a, size_t n) {
    if (n == 0) return 0.f;
    return sqrtf(orc_avx2_dot(a, a, n));
}

/* The reference's in-register exp (avx2.rs:798-866, reused at :900-935 and :1000-1025):
 * clamp to [-87.33655, 88.37626]; k = floor(x*log2e + 0.5); r = x - k*ln2 (one step, mul then
 * sub); degree-6 Taylor by Horner with FMAs; scale by 2^k through the exponent field. */
static inline __m256 exp8_taylor(__m256 x) {
    const __m256 one = _mm256_set1_ps(1.f);
    x = _mm256_max_ps(_mm256_min_ps(x, _mm256_set1_ps(88.37626f)), _mm256_set1_ps(-87.33655f));
    __m256 k = _mm256_floor_ps(_mm256_add_ps(_mm256_mul_ps(x, _mm256_set1_ps(1.44269504088896341f)),
                                             _mm256_set1_ps(0.5f)));
    __m256 r = _mm256_sub_ps(x, _mm256_mul_ps(k, _mm256_set1_ps(0.693147180559945309f)));
    __m256 p = _mm256_set1_ps(0.001388889f);
    p = _mm256_fmadd_ps(p, r, _mm256_set1_ps(0.008333334f));
    p = _mm256_fmadd_ps(p, r, _mm256_set1_ps(0.041666668f));
    p = _mm256_fmadd_ps(p, r, _mm256_set1_ps(0.16666667f));
    p = _mm256_fmadd_ps(p, r, _mm256_set1_ps(0.5f));
    p = _mm256_fmadd_ps(p, r, one);
    p = _mm256_fmadd_ps(p, r, one);
    __m256i e = _mm256_slli_epi32(_mm256_cvtps_epi32(k), 23);
    __m256 two_k = _mm256_castsi256_ps(_mm256_add_epi32(_mm256_castps_si256(one), e));
    return _mm256_mul_ps(p, two_k);
}
ORC_API void orc_avx2_exp(const float* a, float* out, size_t n) {
    size_t i = 0;
    for (; i + 8 <= n; i += 8) _mm256_storeu_ps(out + i, exp8_taylor(_mm256_loadu_ps(a + i)));
    for (; i < n; ++i) out[i] = expf(a[i]);
}
/* avx2.rs:875-949 — 1/(1+exp8(-x)); the <8 tail uses the scalar rule */
ORC_API void orc_avx2_sigmoid(const float* a, float* out, size_t n) {
    const __m256 one = _mm256_set1_ps(1.f);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        __m256 nx = _mm256_sub_ps(_mm256_setzero_ps(), _mm256_loadu_ps(a + i));
        __m256 e = exp8_taylor(nx);
        _mm256_storeu_ps(out + i, _mm256_div_ps(one, _mm256_add_ps(one, e)));
    }
    for (; i < n; ++i) out[i] = sigmoid_libm(a[i]);
}
/* avx2.rs:958-1047 — tanh(u) = (e^{2u}-1)/(e^{2u}+1) with exp8; the <8 tail uses libm tanhf */
ORC_API void orc_avx2_gelu(const float* a, float* out, size_t n) {
    const __m256 one = _mm256_set1_ps(1.f);
    size_t i = 0;
    for (; i + 8 <= n; i += 8) {
        __m256 x = _mm256_loadu_ps(a + i);
        __m256 x3 = _mm256_mul_ps(_mm256_mul_ps(x, x), x);
        __m256 u = _mm256_mul_ps(_mm256_set1_ps(0.7978846f),
                                 _mm256_fmadd_ps(_mm256_set1_ps(0.044715f), x3, x));
        __m256 e = exp8_taylor(_mm256_mul_ps(_mm256_set1_ps(2.f), u));
        __m256 th = _mm256_div_ps(_mm256_sub_ps(e, one), _mm256_add_ps(e, one));
        _mm256_storeu_ps(out + i, _mm256_mul_ps(_mm256_set1_ps(0.5f),
                                                _mm256_mul_ps(x, _mm256_add_ps(one, th))));
    }
    for (; i < n; ++i) out[i] = gelu_libm(a[i]);
}

/* ------------------------------------------------------------------------------------------
 * AVX-512 variants reached only on explicit request and only for dot/sum
 * (src/backends/avx512.rs:151-223).  _mm512_reduce_add_ps folds 16->8->4->2->1 by halves.
 * ---------------------------------------------------------------------------------------- */
__attribute__((target("avx512f"))) static float fold16_add(__m512 v) {
    return _mm512_reduce_add_ps(v);
}
ORC_API int orc_has_avx512(void) { return __builtin_cpu_supports("avx512f"); }

__attribute__((target("avx512f")))
ORC_API float orc_avx512_dot(const float* a, const float* b, size_t n) {
    __m512 v0 = _mm512_setzero_ps(), v1 = v0;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        v0 = _mm512_fmadd_ps(_mm512_loadu_ps(a + i),      _mm512_loadu_ps(b + i),      v0);
        v1 = _mm512_fmadd_ps(_mm512_loadu_ps(a + i + 16), _mm512_loadu_ps(b + i + 16), v1);
    }
    for (; i + 16 <= n; i += 16)
        v0 = _mm512_fmadd_ps(_mm512_loadu_ps(a + i), _mm512_loadu_ps(b + i), v0);
    float r = fold16_add(_mm512_add_ps(v0, v1));
    float tail = 0.f;
    for (; i < n; ++i) { float p = a[i] * b[i]; tail += p; }
    return r + tail;
}
__attribute__((target("avx512f")))
ORC_API float orc_avx512_sum(const float* a, size_t n) {
    __m512 acc = _mm512_setzero_ps();
    size_t i = 0;
    for (; i + 16 <= n; i += 16) acc = _mm512_add_ps(acc, _mm512_loadu_ps(a + i));
    float r = fold16_add(acc);
    float tail = 0.f;
    for (; i < n; ++i) tail += a[i];
    return r + tail;
}

/* ------------------------------------------------------------------------------------------
 * Vector-level entry points: validation + dispatch as in src/vector.rs (default backend AVX2).
 * backend: 0 = scalar, 1 = avx2 (default stamp on x86, src/lib.rs:138-152), 2 = avx512
 * (only dot/sum differ; everything else routes to AVX2 — src/vector.rs:665,713,761,809,2613)
 * ---------------------------------------------------------------------------------------- */
ORC_API int orc_vector_dot(const float* a, size_t na, const float* b, size_t nb, int backend, float* out) {
    if (na != nb) return fail_size(na, nb);                       /* vector.rs:589-594 */
    *out = backend == 0 ? orc_scalar_dot(a, b, na)
         : backend == 2 ? orc_avx512_dot(a, b, na) : orc_avx2_dot(a, b, na);
    return ORC_OK;
}
ORC_API int orc_vector_sum(const float* a, size_t n, int backend, float* out) {
    *out = backend == 0 ? orc_scalar_sum(a, n)                    /* vector.rs:635: empty -> 0 */
         : backend == 2 ? orc_avx512_sum(a, n) : orc_avx2_sum(a, n);
    return ORC_OK;
}
ORC_API int orc_vector_max(const float* a, size_t n, int backend, float* out) {
    if (n == 0) return fail_invalid("Empty vector");              /* vector.rs:654-656 */
    *out = backend == 0 ? orc_scalar_max(a, n) : orc_avx2_max(a, n);
    return ORC_OK;
}
ORC_API int orc_vector_min(const float* a, size_t n, int backend, float* out) {
    if (n == 0) return fail_invalid("Empty vector");              /* vector.rs:702-704 */
    *out = backend == 0 ? orc_scalar_min(a, n) : orc_avx2_min(a, n);
    return ORC_OK;
}
ORC_API int orc_vector_argmax(const float* a, size_t n, int backend, uint64_t* out) {
    if (n == 0) return fail_invalid("Empty vector");              /* vector.rs:750-752 */
    *out = backend == 0 ? orc_scalar_argmax(a, n) : orc_avx2_argmax(a, n);
    return ORC_OK;
}
ORC_API int orc_vector_argmin(const float* a, size_t n, int backend, uint64_t* out) {
    if (n == 0) return fail_invalid("Empty vector");              /* vector.rs:798-800 */
    *out = backend == 0 ? orc_scalar_argmin(a, n) : orc_avx2_argmin(a, n);
    return ORC_OK;
}
ORC_API int orc_vector_norm_l2(const float* a, size_t n, int backend, float* out) {
    *out = n == 0 ? 0.f                                           /* vector.rs:2602-2604 */
         : backend == 0 ? orc_scalar_norm_l2(a, n) : orc_avx2_norm_l2(a, n);
    return ORC_OK;
}
ORC_API int orc_vector_add(const float* a, size_t na, const float* b, size_t nb, float* out) {
    if (na != nb) return fail_size(na, nb);                       /* vector.rs:358-366 */
    orc_avx2_add(a, b, out, na);
    return ORC_OK;
}
ORC_API int orc_vector_mul(const float* a, size_t na, const float* b, size_t nb, float* out) {
    if (na != nb) return fail_size(na, nb);                       /* vector.rs:478-486 */
    orc_avx2_mul(a, b, out, na);
    return ORC_OK;
}
ORC_API int orc_vector_sigmoid(const float* a, size_t n, int backend, float* out) {
    if (n == 0) { snprintf(g_msg, sizeof g_msg, "Empty vector"); return ORC_EMPTY_VECTOR; }
    if (backend == 0) orc_scalar_sigmoid(a, out, n); else orc_avx2_sigmoid(a, out, n);
    return ORC_OK;
}
ORC_API int orc_vector_gelu(const float* a, size_t n, int backend, float* out) {
    if (n == 0) { snprintf(g_msg, sizeof g_msg, "Empty vector"); return ORC_EMPTY_VECTOR; }
    if (backend == 0) orc_scalar_gelu(a, out, n); else orc_avx2_gelu(a, out, n);
    return ORC_OK;
}

/* vector.rs:1540-1553 — max() via the vector's backend, then libm expf per element, a
 * left-to-right f32 sum of the exponentials, one division per element. */
ORC_API int orc_vector_softmax(const float* a, size_t n, int backend, float* out) {
    if (n == 0) { snprintf(g_msg, sizeof g_msg, "Empty vector"); return ORC_EMPTY_VECTOR; }
    float mx = backend == 0 ? orc_scalar_max(a, n) : orc_avx2_max(a, n);
    for (size_t i = 0; i < n; ++i) out[i] = expf(a[i] - mx);
    float s = 0.f;
    for (size_t i = 0; i < n; ++i) s += out[i];
    for (size_t i = 0; i < n; ++i) out[i] = out[i] / s;
    return ORC_OK;
}
/* vector.rs:1605-1623 — out = (x - max) - ln(sum exp(x - max)), evaluated left to right */
ORC_API int orc_vector_log_softmax(const float* a, size_t n, int backend, float* out) {
    if (n == 0) { snprintf(g_msg, sizeof g_msg, "Empty vector"); return ORC_EMPTY_VECTOR; }
    float mx = backend == 0 ? orc_scalar_max(a, n) : orc_avx2_max(a, n);
    float s = 0.f;
    for (size_t i = 0; i < n; ++i) s += expf(a[i] - mx);
    float lse = logf(s);
    for (size_t i = 0; i < n; ++i) { float d = a[i] - mx; out[i] = d - lse; }
    return ORC_OK;
}
/* Row-batched convenience used by the config-5 parity tests: the reference has no batched API,
 * a row is one Vector — so this is literally a loop of the above. */
ORC_API int orc_softmax_rows(const float* a, float* out, size_t rows, size_t cols, int backend, int log_variant) {
    if (cols == 0) { snprintf(g_msg, sizeof g_msg, "Empty vector"); return ORC_EMPTY_VECTOR; }
    #pragma omp parallel for schedule(static)
    for (size_t r = 0; r < rows; ++r) {
        if (log_variant) orc_vector_log_softmax(a + r * cols, cols, backend, out + r * cols);
        else             orc_vector_softmax(a + r * cols, cols, backend, out + r * cols);
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------------------------------
 * Matrix paths (src/matrix.rs).  Row-major everywhere.
 * ---------------------------------------------------------------------------------------- */

/* matrix.rs:1590-1617 — result[j,i] = self[i,j]; blocking does not change values */
ORC_API void orc_transpose(const float* a, float* out, size_t rows, size_t cols) {
    for (size_t ib = 0; ib < rows; ib += 64)
        for (size_t jb = 0; jb < cols; jb += 64) {
            size_t ie = ib + 64 < rows ? ib + 64 : rows, je = jb + 64 < cols ? jb + 64 : cols;
            for (size_t i = ib; i < ie; ++i)
                for (size_t j = jb; j < je; ++j) out[j * rows + i] = a[i * cols + j];
        }
}

/* matrix.rs:572-597 — sum += a*b (mul, then add), k ascending */
ORC_API void orc_matmul_naive(const float* A, const float* B, float* C, size_t m, size_t k, size_t n) {
    for (size_t i = 0; i < m; ++i)
        for (size_t j = 0; j < n; ++j) {
            float s = 0.f;
            for (size_t p = 0; p < k; ++p) { float t = A[i * k + p] * B[p * n + j]; s += t; }
            C[i * n + j] = s;
        }
}

/* matrix.rs:540-569 — rows == 1 fast path: AXPY per k, skipping a_k == 0.0 exactly */
static void matmul_row_vector(const float* a, const float* B, float* c, size_t k, size_t n) {
    memset(c, 0, n * sizeof(float));
    for (size_t p = 0; p < k; ++p) {
        float ak = a[p];
        if (ak == 0.0f) continue;
        const float* brow = B + p * n;
        for (size_t j = 0; j < n; ++j) { float t = ak * brow[j]; c[j] += t; }
    }
}

/* matrix.rs:1476-1536 — per row, 8 running sums over ascending k (mul, then add); scalar tail cols */
static void matmul_tiled8(const float* A, const float* B, float* C, size_t m, size_t k, size_t n) {
    size_t n8 = n / 8 * 8;
    for (size_t i = 0; i < m; ++i) {
        const float* arow = A + i * k;
        for (size_t j0 = 0; j0 < n8; j0 += 8) {
            float acc[8] = {0};
            for (size_t p = 0; p < k; ++p) {
                float av = arow[p];
                const float* b = B + p * n + j0;
                for (int l = 0; l < 8; ++l) { float t = av * b[l]; acc[l] += t; }
            }
            memcpy(C + i * n + j0, acc, sizeof acc);
        }
        for (size_t j = n8; j < n; ++j) {
            float s = 0.f;
            for (size_t p = 0; p < k; ++p) { float t = arow[p] * B[p * n + j]; s += t; }
            C[i * n + j] = s;
        }
    }
}

/* matrix.rs:674-688 — lo128+hi128 then two hadds: ((s0+s1)+(s2+s3)) */
static inline float fold8_hadd(__m256 v) {
    __m128 s = _mm_add_ps(_mm256_castps256_ps128(v), _mm256_extractf128_ps(v, 1));
    s = _mm_hadd_ps(s, s);
    s = _mm_hadd_ps(s, s);
    return _mm_cvtss_f32(s);
}

/* matrix.rs:615-672 — 4 rows x 1 column over one k-block: 8-wide FMA chains, hadd fold,
 * then the (<8) remainder as separate mul + add onto the folded value */
static inline void microkernel_4x1(const float* a0, const float* a1, const float* a2, const float* a3,
                                   const float* bt, size_t len, float out[4]) {
    __m256 c0 = _mm256_setzero_ps(), c1 = c0, c2 = c0, c3 = c0;
    size_t full = len / 8;
    for (size_t q = 0; q < full; ++q) {
        __m256 bv = _mm256_loadu_ps(bt + 8 * q);
        c0 = _mm256_fmadd_ps(_mm256_loadu_ps(a0 + 8 * q), bv, c0);
        c1 = _mm256_fmadd_ps(_mm256_loadu_ps(a1 + 8 * q), bv, c1);
        c2 = _mm256_fmadd_ps(_mm256_loadu_ps(a2 + 8 * q), bv, c2);
        c3 = _mm256_fmadd_ps(_mm256_loadu_ps(a3 + 8 * q), bv, c3);
    }
    out[0] = fold8_hadd(c0); out[1] = fold8_hadd(c1); out[2] = fold8_hadd(c2); out[3] = fold8_hadd(c3);
    for (size_t p = 8 * full; p < len; ++p) {
        float t;
        t = a0[p] * bt[p]; out[0] += t;
        t = a1[p] * bt[p]; out[1] += t;
        t = a2[p] * bt[p]; out[2] += t;
        t = a3[p] * bt[p]; out[3] += t;
    }
}

/* One 64x64x64 (edge-clipped) L2 block: rows in groups of 4 through the microkernel, leftover
 * rows through Avx2Backend::dot on the k-block; every partial is ADDED into C (matrix.rs:1040-1098). */
static void l2_block(const float* A, const float* Bt, float* C, size_t k, size_t n,
                     size_t i0, size_t i1, size_t j0, size_t j1, size_t p0, size_t p1) {
    size_t len = p1 - p0, i = i0;
    for (; i + 4 <= i1; i += 4) {
        const float* r0 = A + i * k + p0;
        for (size_t j = j0; j < j1; ++j) {
            float part[4];
            microkernel_4x1(r0, r0 + k, r0 + 2 * k, r0 + 3 * k, Bt + j * k + p0, len, part);
            C[i * n + j] += part[0];
            C[(i + 1) * n + j] += part[1];
            C[(i + 2) * n + j] += part[2];
            C[(i + 3) * n + j] += part[3];
        }
    }
    for (; i < i1; ++i)
        for (size_t j = j0; j < j1; ++j)
            C[i * n + j] += orc_avx2_dot(A + i * k + p0, Bt + j * k + p0, len);
}

static inline size_t zmin(size_t a, size_t b) { return a < b ? a : b; }

/* rows [r0,r1) of the 3-level blocked product (matrix.rs:711-905 / :1015-1098): L3 = 256, L2 = 64;
 * for any C[i,j] the k-blocks arrive in ascending order. */
static void l3_rows(const float* A, const float* Bt, float* C, size_t k, size_t n, size_t r0, size_t r1) {
    for (size_t J = 0; J < n; J += 256) {
        size_t Je = zmin(J + 256, n);
        for (size_t P = 0; P < k; P += 256) {
            size_t Pe = zmin(P + 256, k);
            for (size_t i = r0; i < r1; i += 64)
                for (size_t j = J; j < Je; j += 64)
                    for (size_t p = P; p < Pe; p += 64)
                        l2_block(A, Bt, C, k, n, i, zmin(i + 64, r1), j, zmin(j + 64, Je), p, zmin(p + 64, Pe));
        }
    }
}

/* matrix.rs:912-1401.  parallel != 0 mirrors `--features parallel` (rayon over 256-row blocks,
 * only when all dims >= 1024, :942-1011) with OpenMP; per-element results are identical either way. */
static void matmul_blocked(const float* A, const float* B, float* C, size_t m, size_t k, size_t n, int parallel) {
    if (m <= 32 || k <= 32 || n <= 32) {
        /* matmul_simd_simple (matrix.rs:1407-1465): transpose B, one Avx2 dot per element */
        float* Bt = (float*)malloc(sizeof(float) * k * n + 64);
        orc_transpose(B, Bt, k, n);
        for (size_t i = 0; i < m; ++i)
            for (size_t j = 0; j < n; ++j) C[i * n + j] = orc_avx2_dot(A + i * k, Bt + j * k, k);
        free(Bt);
        return;
    }
    float* Bt = (float*)malloc(sizeof(float) * k * n + 64);
    orc_transpose(B, Bt, k, n);
    memset(C, 0, sizeof(float) * m * n);
    if (m >= 512 && k >= 512 && n >= 512) {
        int par = parallel && m >= 1024 && k >= 1024 && n >= 1024;
        size_t nblk = (m + 255) / 256;
        #pragma omp parallel for schedule(dynamic) if (par)
        for (size_t b = 0; b < nblk; ++b) l3_rows(A, Bt, C, k, n, b * 256, zmin(b * 256 + 256, m));
    } else {
        /* 2-level blocking (matrix.rs:1227-1300) */
        for (size_t i = 0; i < m; i += 64)
            for (size_t j = 0; j < n; j += 64)
                for (size_t p = 0; p < k; p += 64)
                    l2_block(A, Bt, C, k, n, i, zmin(i + 64, m), j, zmin(j + 64, n), p, zmin(p + 64, k));
    }
    free(Bt);
}

/* Matrix::matmul routing (matrix.rs:285-358), default features (no gpu), AVX2-stamped matrices */
static void matmul_route(const float* A, const float* B, float* C, size_t m, size_t k, size_t n, int parallel) {
    if (m == 1) { matmul_row_vector(A, B, C, k, n); return; }
    if (m >= 64 || k >= 64 || n >= 64) {
        size_t mx = m > k ? m : k; if (n > mx) mx = n;
        if (mx < 512) matmul_tiled8(A, B, C, m, k, n);
        else matmul_blocked(A, B, C, m, k, n, parallel);
    } else {
        orc_matmul_naive(A, B, C, m, k, n);
    }
}

ORC_API int orc_matmul(const float* A, size_t a_rows, size_t a_cols,
                       const float* B, size_t b_rows, size_t b_cols, float* C, int parallel) {
    if (a_cols != b_rows) {
        snprintf(g_msg, sizeof g_msg,
                 "Matrix dimension mismatch for multiplication: %zu\xC3\x97%zu \xC3\x97 %zu\xC3\x97%zu "
                 "(inner dimensions %zu and %zu must match)",
                 a_rows, a_cols, b_rows, b_cols, a_cols, b_rows);
        return ORC_INVALID_INPUT;
    }
    matmul_route(A, B, C, a_rows, a_cols, b_cols, parallel);
    return ORC_OK;
}
/* forced-path entry for tests that compare paths (reference does the same: matrix.rs:2285-2715) */
ORC_API void orc_matmul_simd(const float* A, const float* B, float* C, size_t m, size_t k, size_t n, int parallel) {
    matmul_blocked(A, B, C, m, k, n, parallel);
}

/* matrix.rs:383-441 */
ORC_API int orc_batched_matmul(const float* A, size_t a_len, const float* B, size_t b_len, float* C,
                               size_t batch, size_t m, size_t k, size_t n, int parallel) {
    if (a_len != batch * m * k) {
        snprintf(g_msg, sizeof g_msg, "A data size mismatch: expected %zu (%zu\xC3\x97%zu\xC3\x97%zu), got %zu",
                 batch * m * k, batch, m, k, a_len);
        return ORC_INVALID_INPUT;
    }
    if (b_len != batch * k * n) {
        snprintf(g_msg, sizeof g_msg, "B data size mismatch: expected %zu (%zu\xC3\x97%zu\xC3\x97%zu), got %zu",
                 batch * k * n, batch, k, n, b_len);
        return ORC_INVALID_INPUT;
    }
    for (size_t b = 0; b < batch; ++b)
        matmul_route(A + b * m * k, B + b * k * n, C + b * m * n, m, k, n, parallel);
    return ORC_OK;
}
/* matrix.rs:464-527 — sequential loop over batch*heads; `parallel` here fans heads out over
 * OpenMP threads for the all-cores CPU baseline only (values per head are unchanged) */
ORC_API int orc_batched_matmul_4d(const float* A, size_t a_len, const float* B, size_t b_len, float* C,
                                  size_t batch, size_t heads, size_t m, size_t k, size_t n, int parallel) {
    size_t total = batch * heads;
    if (a_len != total * m * k) {
        snprintf(g_msg, sizeof g_msg,
                 "A data size mismatch: expected %zu (%zu\xC3\x97%zu\xC3\x97%zu\xC3\x97%zu), got %zu",
                 total * m * k, batch, heads, m, k, a_len);
        return ORC_INVALID_INPUT;
    }
    if (b_len != total * k * n) {
        snprintf(g_msg, sizeof g_msg,
                 "B data size mismatch: expected %zu (%zu\xC3\x97%zu\xC3\x97%zu\xC3\x97%zu), got %zu",
                 total * k * n, batch, heads, k, n, b_len);
        return ORC_INVALID_INPUT;
    }
    #pragma omp parallel for schedule(dynamic) if (parallel)
    for (size_t h = 0; h < total; ++h)
        matmul_route(A + h * m * k, B + h * k * n, C + h * m * n, m, k, n, 0);
    return ORC_OK;
}

/* matrix.rs:1657-1741 — one Avx2 dot per row; rayon over rows at >= 4096 rows with `parallel` */
ORC_API int orc_matvec(const float* A, size_t rows, size_t cols, const float* v, size_t vlen, float* y, int parallel) {
    if (vlen != cols) {
        snprintf(g_msg, sizeof g_msg,
                 "Vector length %zu does not match matrix columns %zu for matrix-vector multiplication",
                 vlen, cols);
        return ORC_INVALID_INPUT;
    }
    int par = parallel && rows >= 4096;
    #pragma omp parallel for schedule(static) if (par)
    for (size_t i = 0; i < rows; ++i) y[i] = orc_avx2_dot(A + i * cols, v, cols);
    return ORC_OK;
}

/* `parallel`-feature maps (vector.rs:369-390): rayon chunks of 65 536 at >= 100 000 elements.
 * op: 0 add, 1 mul, 2 sigmoid, 3 gelu (sigmoid/gelu have no rayon path in the reference; the
 * chunked form is used for the all-cores baseline only and gives identical values because
 * 65 536 is a multiple of the 8-lane width). */
ORC_API void orc_map_parallel(int op, const float* a, const float* b, float* out, size_t n) {
    const size_t chunk = 65536;
    size_t nchunks = (n + chunk - 1) / chunk;
    #pragma omp parallel for schedule(static)
    for (size_t c = 0; c < nchunks; ++c) {
        size_t o = c * chunk, len = zmin(chunk, n - o);
        switch (op) {
            case 0: orc_avx2_add(a + o, b + o, out + o, len); break;
            case 1: orc_avx2_mul(a + o, b + o, out + o, len); break;
            case 2: orc_avx2_sigmoid(a + o, out + o, len); break;
            default: orc_avx2_gelu(a + o, out + o, len); break;
        }
    }
}

/* ------------------------------------------------------------------------------------------
 * f64 "truth" helpers for condition-aware tolerances: these are NOT reference
 * behaviour, they are the yardstick both the reference order and the GPU result are measured on.
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_f64_sum(const float* a, size_t n, double* sum, double* abs_sum) {
    double s = 0, t = 0;
    #pragma omp parallel for reduction(+:s,t) schedule(static)
    for (size_t i = 0; i < n; ++i) { s += (double)a[i]; t += fabs((double)a[i]); }
    *sum = s; *abs_sum = t;
}
ORC_API void orc_f64_dot(const float* a, const float* b, size_t n, double* dot, double* abs_dot) {
    double s = 0, t = 0;
    #pragma omp parallel for reduction(+:s,t) schedule(static)
    for (size_t i = 0; i < n; ++i) { double p = (double)a[i] * (double)b[i]; s += p; t += fabs(p); }
    *dot = s; *abs_dot = t;
}
/* C[i,j] in f64 for a list of sampled (i,j) pairs, plus sum |a||b| for the tolerance scale */
ORC_API void orc_f64_matmul_samples(const float* A, const float* B, size_t k, size_t n,
                                    const uint64_t* rows, const uint64_t* cols, size_t count,
                                    double* out, double* out_abs) {
    #pragma omp parallel for schedule(static)
    for (size_t s = 0; s < count; ++s) {
        const float* a = A + rows[s] * k; const float* b = B + cols[s];
        double acc = 0, aacc = 0;
        for (size_t p = 0; p < k; ++p) { double t = (double)a[p] * (double)b[p * n]; acc += t; aacc += fabs(t); }
        out[s] = acc; out_abs[s] = aacc;
    }
}
ORC_API int orc_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
ORC_API void orc_set_threads(int t) {
#ifdef _OPENMP
    omp_set_num_threads(t);
#else
    (void)t;
#endif
}

/* ------------------------------------------------------------------------------------------
 * Remaining VectorBackend surface, scalar backend: one restatement per
 * function of src/backends/scalar.rs, same operation order, libm for transcendentals.
 * op codes are private to the oracle (oracle/__init__.py holds the name table).
 * ---------------------------------------------------------------------------------------- */
/* scalar.rs:32-60 (sub, div), :250-260 (scale), :262-266 (abs), :268-272 (clamp: val.max(min).min(max)),
 * :274-279 (lerp: a + t*(b-a)), :281-286 (fma: a*b + c, NOT fused), :288-294 (relu), :296-300 (exp),
 * :342-353 (swish), :355-480 (tanh, sqrt, recip, ln, log2, log10, sin, cos, tan, floor, ceil, round) */
ORC_API void orc_scalar_map(int op, const float* a, const float* b, const float* c, float p0, float p1,
                            float* out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        const float x = a[i];
        float r;
        switch (op) {
            case 0: r = x - b[i]; break;                                  /* sub */
            case 1: r = x / b[i]; break;                                  /* div */
            case 2: r = x * p0; break;                                    /* scale */
            case 3: r = fabsf(x); break;                                  /* abs */
            case 4: r = fminf(fmaxf(x, p0), p1); break;                   /* clamp (f32::max/min ignore NaN like fmaxf/fminf) */
            case 5: { float d = b[i] - x; float t = p0 * d; r = x + t; } break;   /* lerp */
            case 6: { float m = x * b[i]; r = m + c[i]; } break;          /* fma (unfused; -ffp-contract=off) */
            case 7: r = x > 0.0f ? x : 0.0f; break;                       /* relu */
            case 8: r = expf(x); break;
            case 9: r = x < -50.0f ? 0.0f : (x > 50.0f ? x : x * (1.0f / (1.0f + expf(-x)))); break;   /* swish */
            case 10: r = tanhf(x); break;
            case 11: r = sqrtf(x); break;
            case 12: r = 1.0f / x; break;                                 /* recip */
            case 13: r = logf(x); break;
            case 14: r = log2f(x); break;
            case 15: r = log10f(x); break;
            case 16: r = sinf(x); break;
            case 17: r = cosf(x); break;
            case 18: r = tanf(x); break;
            case 19: r = floorf(x); break;
            case 20: r = ceilf(x); break;
            case 21: r = roundf(x); break;                                /* half away from zero == f32::round */
            default: r = x;
        }
        out[i] = r;
    }
}

/* ------------------------------------------------------------------------------------------
 * The rest of Vector's element-wise / statistics API: the scalar closures in src/vector.rs, same operation
 * order, libm (Rust std f32 methods) for transcendentals.  op codes are private to the oracle.
 * Rust std semantics restated: f32::signum (+-1 by sign bit, NaN -> NaN), f32::fract = x - x.trunc(),
 * f32::min / max (the non-NaN operand, like fminf / fmaxf), f32::copysign, f32::powf.
 * f32::asinh / acosh / atanh are computed inside Rust std by formulas that vary between std versions (not under
 * /root/reference); the oracle uses glibc's and the parity tests hold both sides to an ulp bound against f64.
 * ---------------------------------------------------------------------------------------- */
ORC_API void orc_vector_map(int op, const float* a, const float* b, float p0, float p1, float* out, size_t n) {
    /* src/vector.rs:2546-2570: LAMBDA * ALPHA * (exp(x) - 1) parses as (LAMBDA * ALPHA) * (...), constants folded in f32 */
    const float kLambda = 1.0507009873554804934193349852946f;
    const float kAlpha = 1.6732632423543772848170429916717f;
    const float kLA = kLambda * kAlpha;
    for (size_t i = 0; i < n; ++i) {
        const float x = a[i];
        float r;
        switch (op) {
            case 0: r = -x; break;                                                    /* neg :4399 */
            case 1: r = isnan(x) ? x : copysignf(1.0f, x); break;                     /* signum :4261 */
            case 2: r = truncf(x); break;                                             /* trunc :4211 */
            case 3: r = x - truncf(x); break;                                         /* fract :4237 */
            case 4: r = sinhf(x); break;                                              /* :3885 */
            case 5: r = coshf(x); break;                                              /* :3916 */
            case 6: r = asinf(x); break;                                              /* :3761 */
            case 7: r = acosf(x); break;                                              /* :3807 */
            case 8: r = atanf(x); break;                                              /* :3855 */
            case 9: r = asinhf(x); break;                                             /* :4059 */
            case 10: r = acoshf(x); break;                                            /* :4089 */
            case 11: r = atanhf(x); break;                                            /* :4111 */
            case 12: r = x <= -3.0f ? 0.0f : (x >= 3.0f ? x : x * (x + 3.0f) / 6.0f); break;   /* hardswish :2409-2432 */
            case 13:                                                                  /* mish :2477-2500 */
                if (x < -20.0f) r = 0.0f;
                else if (x > 20.0f) r = x;
                else { float sp = logf(1.0f + expf(x)); r = x * tanhf(sp); }
                break;
            case 14: r = x > 0.0f ? kLambda * x : kLA * (expf(x) - 1.0f); break;      /* selu :2546-2570 */
            case 15: r = x > 0.0f ? x : p0 * x; break;                                /* leaky_relu :2014-2019 */
            case 16: r = x > 0.0f ? x : p0 * (expf(x) - 1.0f); break;                 /* elu :2118-2122 */
            case 17: r = powf(x, p0); break;                                          /* pow :3342 */
            case 18: r = fminf(fmaxf(x, p0), p1); break;                              /* clip :1475-1480 */
            case 19: r = fminf(x, b[i]); break;                                       /* minimum :4328 */
            case 20: r = fmaxf(x, b[i]); break;                                       /* maximum :4364 */
            case 21: r = copysignf(x, b[i]); break;                                   /* copysign :4292 */
            case 22: r = (x - p0) * p1; break;                                        /* (x - mean) * inv_std :1195-1200 */
            default: r = x;
        }
        out[i] = r;
    }
}

/* Vector::layer_norm_simple (src/vector.rs:1386-1412): mean = sum / n (`sum` is the caller's Vector::sum — passed in so
 * the restatement can use the backend under test), sequential f32 variance, (x - mean) * inv_std */
ORC_API void orc_layer_norm_simple(const float* x, float sum, float eps, float* y, size_t n) {
    const float mean = sum / (float)n;
    float var = 0.f;
    for (size_t i = 0; i < n; ++i) { float d = x[i] - mean; var += d * d; }
    var = var / (float)n;
    const float inv_std = 1.0f / sqrtf(var + eps);
    for (size_t i = 0; i < n; ++i) y[i] = (x[i] - mean) * inv_std;
}

/* Matrix::embedding_lookup (src/matrix.rs:2008-2041): validate every index first, then copy one row per index.
 * Returns 0, or 2 (InvalidInput) with the reference's message in orc_last_message(). */
ORC_API int orc_embedding_lookup(const float* table, size_t rows, size_t cols, const uint64_t* idx, size_t n, float* out) {
    for (size_t i = 0; i < n; ++i)
        if (idx[i] >= rows) {
            snprintf(g_msg, sizeof g_msg, "Index %llu at position %zu is out of bounds for embedding table with %zu rows",
                     (unsigned long long)idx[i], i, rows);
            return ORC_INVALID_INPUT;
        }
    for (size_t r = 0; r < n; ++r) memcpy(out + r * cols, table + idx[r] * cols, cols * sizeof(float));
    return 0;
}

/* scalar.rs:170-183 — Kahan-compensated sequential sum */
ORC_API float orc_scalar_sum_kahan(const float* a, size_t n) {
    float sum = 0.f, c = 0.f;
    for (size_t i = 0; i < n; ++i) {
        float y = a[i] - c;
        float t = sum + y;
        c = (t - sum) - y;
        sum = t;
    }
    return sum;
}
/* scalar.rs:218-228 */
ORC_API float orc_scalar_norm_l1(const float* a, size_t n) {
    float s = 0.f;
    for (size_t i = 0; i < n; ++i) s += fabsf(a[i]);
    return s;
}
/* scalar.rs:234-247 — starts at 0.0, strict compare: a NaN never wins */
ORC_API float orc_scalar_norm_linf(const float* a, size_t n) {
    float m = 0.f;
    for (size_t i = 0; i < n; ++i) {
        float v = fabsf(a[i]);
        if (v > m) m = v;
    }
    return m;
}

/* ------------------------------------------------------------------------------------------
 * Callers next to the path
 * ---------------------------------------------------------------------------------------- */
/* src/matrix.rs:1782-1816 — result = sum_i row_i.scale(v[i]) accumulated in order: unfused mul then add */
ORC_API void orc_vecmat(const float* v, const float* a, size_t rows, size_t cols, float* y) {
    for (size_t j = 0; j < cols; ++j) y[j] = 0.f;
    for (size_t i = 0; i < rows; ++i)
        for (size_t j = 0; j < cols; ++j) {
            float s = a[i * cols + j] * v[i];
            y[j] = y[j] + s;
        }
}
/* src/vector.rs:1316-1362 — mean = sum/n (sequential f32 sum here: the scalar backend's sum), variance =
 * sequential sum of (x-mean)^2 / n, inv_std = 1/sqrt(var+eps), y = g*(x-mean)*inv_std + b (unfused) */
ORC_API void orc_layer_norm(const float* x, const float* g, const float* b, float eps, float* y, size_t n) {
    float sum = 0.f;
    for (size_t i = 0; i < n; ++i) sum += x[i];
    const float mean = sum / (float)n;
    float var = 0.f;
    for (size_t i = 0; i < n; ++i) { float d = x[i] - mean; var += d * d; }
    var = var / (float)n;
    const float inv_std = 1.0f / sqrtf(var + eps);
    for (size_t i = 0; i < n; ++i) {
        float t = g[i] * (x[i] - mean);
        t = t * inv_std;
        y[i] = t + b[i];
    }
}

/* src/matrix.rs:1868-1950 — valid-padding cross-correlation, scalar loop: sum += input * kernel (unfused) */
ORC_API void orc_convolve2d(const float* in, size_t rows, size_t cols, const float* k, size_t kr, size_t kc, float* out) {
    const size_t orows = rows - kr + 1, ocols = cols - kc + 1;
    for (size_t r = 0; r < orows; ++r)
        for (size_t c = 0; c < ocols; ++c) {
            float sum = 0.f;
            for (size_t a = 0; a < kr; ++a)
                for (size_t b = 0; b < kc; ++b) {
                    float p = in[(r + a) * cols + (c + b)] * k[a * kc + b];
                    sum = sum + p;
                }
            out[r * ocols + c] = sum;
        }
}

/* src/eigen.rs:143-306 — SymmetricEigen::compute_jacobi: cyclic Jacobi, row-cyclic pair order, rotations applied one
 * after the other (src/eigen.rs:248-306: tau, t, c, s by the Golub & Van Loan formulas; a_pp -= t*a_pq, a_qq += t*a_pq,
 * a_pq = 0; columns p, q of A mirrored into rows; columns p, q of V), threshold CONVERGENCE_THRESHOLD (1e-7) *
 * max(||A||_F, 1) (src/eigen.rs:150-151), at most MAX_JACOBI_SWEEPS = 50 sweeps; eigenvalues sorted descending with
 * a stable sort and the eigenvector COLUMNS permuted alike (src/eigen.rs:183-203).
 * Returns 0, or 2 when it does not converge ("Jacobi algorithm failed to converge after 50 sweeps"). */
ORC_API int orc_symmetric_eigen(const float* matrix, size_t n, float* eigenvalues, float* eigenvectors) {
    float* a = (float*)malloc(sizeof(float) * n * n);
    float* v = (float*)calloc(n * n, sizeof(float));
    size_t* idx = (size_t*)malloc(sizeof(size_t) * n);
    memcpy(a, matrix, sizeof(float) * n * n);
    float frob = 0.f;
    for (size_t i = 0; i < n * n; ++i) { float sq = a[i] * a[i]; frob = frob + sq; }
    float root = sqrtf(frob);
    const float tolerance = 1e-7f * (root > 1.0f ? root : 1.0f);
    for (size_t i = 0; i < n; ++i) v[i * n + i] = 1.0f;
    int status = 2;
    for (int sweep = 0; sweep < 50 && status != 0; ++sweep) {
        int converged = 1;
        for (size_t p = 0; p < n; ++p)
            for (size_t q = p + 1; q < n; ++q) {
                if (fabsf(a[p * n + q]) < tolerance) continue;
                converged = 0;
                const float app = a[p * n + p], aqq = a[q * n + q], apq = a[p * n + q];
                if (fabsf(apq) < 1e-15f) continue;
                const float two_apq = 2.0f * apq;
                const float tau = (aqq - app) / two_apq;
                float tt = tau * tau;
                float root1 = sqrtf(1.0f + tt);
                float t = tau >= 0.0f ? 1.0f / (tau + root1) : -1.0f / (-tau + root1);
                float t2 = t * t;
                const float c = 1.0f / sqrtf(1.0f + t2);
                const float s = t * c;
                float tapq = t * apq;
                a[p * n + p] = app - tapq;
                a[q * n + q] = aqq + tapq;
                a[p * n + q] = 0.0f;
                a[q * n + p] = 0.0f;
                for (size_t k = 0; k < n; ++k) {
                    if (k == p || k == q) continue;
                    const float akp = a[k * n + p], akq = a[k * n + q];
                    float x0 = c * akp, x1 = s * akq, y0 = s * akp, y1 = c * akq;
                    a[k * n + p] = x0 - x1;
                    a[p * n + k] = a[k * n + p];
                    a[k * n + q] = y0 + y1;
                    a[q * n + k] = a[k * n + q];
                }
                for (size_t k = 0; k < n; ++k) {
                    const float vkp = v[k * n + p], vkq = v[k * n + q];
                    float x0 = c * vkp, x1 = s * vkq, y0 = s * vkp, y1 = c * vkq;
                    v[k * n + p] = x0 - x1;
                    v[k * n + q] = y0 + y1;
                }
            }
        if (converged) status = 0;
    }
    if (status == 0) {
        /* stable descending sort of the diagonal (insertion sort == any stable sort) */
        for (size_t i = 0; i < n; ++i) idx[i] = i;
        for (size_t i = 1; i < n; ++i) {
            size_t key = idx[i];
            size_t j = i;
            while (j > 0 && a[idx[j - 1] * n + idx[j - 1]] < a[key * n + key]) { idx[j] = idx[j - 1]; --j; }
            idx[j] = key;
        }
        for (size_t c2 = 0; c2 < n; ++c2) {
            eigenvalues[c2] = a[idx[c2] * n + idx[c2]];
            for (size_t r = 0; r < n; ++r) eigenvectors[r * n + c2] = v[r * n + idx[c2]];
        }
    }
    free(a); free(v); free(idx);
    return status;
}
