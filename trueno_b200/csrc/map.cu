// map.cu — elementwise kernels: the whole `VectorBackend` map surface (src/backends/mod.rs:52-385) and the rest of
// Vector's element-wise API (neg ... atanh, hardswish, mish, selu, leaky_relu, elu, pow, minimum, maximum, copysign;
// scalar closures in src/vector.rs:1448-4410).
//
// Replaces Avx2Backend::{add,sub,mul,div,scale,abs,clamp,lerp,fma,relu,exp,sigmoid,gelu,swish,tanh,sqrt,
// recip,ln,log2,log10,sin,cos,tan,floor,ceil,round} (src/backends/avx2.rs) with the SEMANTICS of the scalar
// backend (src/backends/scalar.rs:20-60, 266-480).  Single-operation maps (add, sub, mul, div, scale, abs,
// clamp, relu, sqrt, recip, floor, ceil, round) and the unfused two-operation maps (lerp = a + t*(b-a),
// fma = a*b + c — Rust never contracts them, scalar.rs:271, :283) are spelled with explicitly rounded
// intrinsics, so they match the reference bit for bit.  Transcendentals use CUDA's accurate libm-class
// functions (expf, logf, sinf, ... <= 2 ulp; tanf <= 4 ulp) — no fast-math intrinsics on the parity path.
// sigmoid / swish keep the scalar backend's +-50 cut-offs as selects; gelu evaluates the reference's
// u = k*(x + c*x^3) in its operation order and then the exact identity 0.5*(1 + tanh u) = 1/(1 + e^(-2u)).
//
// HBM-bound streaming: 128-bit loads/stores, ONE 8 KiB tile per CTA on a flat (non-persistent) grid.
// Measured on B200 (scripts/exp/exp_map.cu, 131 M elements): flat grids reach 6.6-6.9 TB/s where a
// persistent one-wave grid-stride loop of the same body reaches 5.0-6.5 TB/s — the block scheduler
// keeps every SM's load queue full while persistent CTAs drift into lock-step load/compute phases.
// Algorithmic bytes per element: 4 B per input + 4 B output (unary 8 B, binary 12 B, fma 16 B).
#include "common.cuh"

namespace trn {

constexpr int kThreads = 256;
constexpr int kUnroll = 2;   // float4 per thread per input: 8 KiB tiles

__host__ __device__ constexpr int map_arity(Map op) {
    return (op == Map::Add || op == Map::Sub || op == Map::Mul || op == Map::Div || op == Map::Lerp ||
            op == Map::Minimum || op == Map::Maximum || op == Map::Copysign) ? 2
           : op == Map::Fma ? 3 : 1;
}

template <Map OP>
__device__ __forceinline__ float apply(float x, float y, float z, float p0, float p1) {
    switch (OP) {
        case Map::Add: return __fadd_rn(x, y);
        case Map::Sub: return __fsub_rn(x, y);
        case Map::Mul: return __fmul_rn(x, y);
        case Map::Div: return __fdiv_rn(x, y);
        case Map::Scale: return __fmul_rn(x, p0);
        case Map::Abs: return fabsf(x);
        case Map::Clamp: return fminf(fmaxf(x, p0), p1);                 // val.max(min).min(max): NaN -> min
        case Map::Lerp: return __fadd_rn(x, __fmul_rn(p0, __fsub_rn(y, x)));   // a + t*(b - a), unfused
        case Map::Fma: return __fadd_rn(__fmul_rn(x, y), z);             // a*b + c, unfused (scalar.rs:283)
        case Map::Relu: return x > 0.0f ? x : 0.0f;                       // NaN and -0.0 -> +0.0 (scalar.rs:293)
        case Map::Exp: return expf(x);
        case Map::Sigmoid: {
            // src/backends/scalar.rs:313-324 (cut-offs applied as selects: no divergence)
            const float s = 1.0f / (1.0f + expf(-x));
            return x < -50.0f ? 0.0f : (x > 50.0f ? 1.0f : s);
        }
        case Map::Gelu: {
            // src/backends/scalar.rs:330-340: 0.5*x*(1 + tanh(u)), u = k*(x + c*x^3) with the reference's
            // operation order for u, then 0.5*(1 + tanh(u)) = 1/(1 + e^(-2u)): branch-free and WITHOUT the
            // 1 + tanh cancellation the reference's form has for x << 0 (tolerance: 4 ulp + 4*2^-24*|x|).
            const float x3 = __fmul_rn(__fmul_rn(x, x), x);
            const float u = __fmul_rn(0.7978846f, __fadd_rn(x, __fmul_rn(0.044715f, x3)));
            // __fdividef (x * rcp.approx(d), <= 2 ulp), not x / d: the IEEE division (Newton steps + range fix-ups) made this
            // map ISSUE-bound (ncu: issue slots 84 % busy, DRAM 74 % against sigmoid's 71 % / 81 %; 161.5 us at 131 M elements,
            // 157.7 with x * (1 / d), 151.7 = sigmoid's time with this).  Measured against the reference's expression in f64
            // over 131 M samples each of N(0,1)*4, U[-12,12], U[-3,3] (scripts/exp/exp_gelu.py): worst error 0.28 of the
            // stated bound 4 ulp(y) + 4 * 2^-24 |x|, 3.4 ulp(y) on [-3, 3].  d = inf (x < -10) gives -0 as the reference does;
            // d in (2^126, 2^128), where __fdividef returns 0, is a result of magnitude < 1e-37.
            return __fdividef(x, 1.0f + expf(-2.0f * u));
        }
        case Map::Swish: {
            // src/backends/scalar.rs:342-353: x * sigmoid(x), x < -50 -> 0, x > 50 -> x
            const float s = __fmul_rn(x, 1.0f / (1.0f + expf(-x)));
            return x < -50.0f ? 0.0f : (x > 50.0f ? x : s);
        }
        case Map::Tanh: return tanhf(x);
        case Map::Sqrt: return sqrtf(x);
        case Map::Recip: return __frcp_rn(x);
        case Map::Ln: return logf(x);
        case Map::Log2: return log2f(x);
        case Map::Log10: return log10f(x);
        case Map::Sin: return sinf(x);
        case Map::Cos: return cosf(x);
        case Map::Tan: return tanf(x);
        case Map::Floor: return floorf(x);
        case Map::Ceil: return ceilf(x);
        case Map::Round: return roundf(x);                                // half away from zero == f32::round
        // ---- the rest of Vector's element-wise API: scalar closures in src/vector.rs, Rust f32 semantics
        case Map::Neg: return -x;                                         // :4399 (sign flip, NaN payload kept)
        case Map::Signum: return x != x ? x : copysignf(1.0f, x);         // :4261 f32::signum: +-1 by sign bit, NaN -> NaN
        case Map::Trunc: return truncf(x);                                // :4211
        case Map::Fract: return __fsub_rn(x, truncf(x));                  // :4237 f32::fract = x - x.trunc()
        case Map::Sinh: return sinhf(x);                                  // :3885
        case Map::Cosh: return coshf(x);                                  // :3916
        case Map::Asin: return asinf(x);                                  // :3761
        case Map::Acos: return acosf(x);                                  // :3807
        case Map::Atan: return atanf(x);                                  // :3855
        case Map::Asinh: return asinhf(x);                                // :4059
        case Map::Acosh: return acoshf(x);                                // :4089
        case Map::Atanh: return atanhf(x);                                // :4111
        case Map::Hardswish: {                                            // :2409-2432, x*(x+3)/6 unfused -> bit-exact
            const float mid = __fdiv_rn(__fmul_rn(x, __fadd_rn(x, 3.0f)), 6.0f);
            return x <= -3.0f ? 0.0f : (x >= 3.0f ? x : mid);
        }
        case Map::Mish: {
            // :2477-2500: x * tanh(ln(1 + e^x)), cut-offs +-20.  With n = e^x the exact identity
            // tanh(ln(1 + n)) = ((1+n)^2 - 1) / ((1+n)^2 + 1) = w / (w + 2), w = n (n + 2), needs ONE exponential and a
            // division instead of exp + log + tanh (ncu: 90 % issue-bound, 4.4 TB/s -> HBM-bound), has no cancellation
            // (all terms positive; n <= e^20, w <= 2.4e17), and drops the reference's own 1 + e^x rounding for x << 0.
            const float n = expf(x);
            const float w = n * (n + 2.0f);
            const float r = x * (w / (w + 2.0f));
            return x < -20.0f ? 0.0f : (x > 20.0f ? x : r);
        }
        case Map::Selu: {                                                 // :2546-2570; LAMBDA * ALPHA folds to one f32 constant
            constexpr float kLambda = 1.0507009873554804934193349852946f;
            constexpr float kAlpha = 1.6732632423543772848170429916717f;
            constexpr float kLA = kLambda * kAlpha;
            return x > 0.0f ? __fmul_rn(kLambda, x) : __fmul_rn(kLA, __fsub_rn(expf(x), 1.0f));
        }
        case Map::LeakyRelu: return x > 0.0f ? x : __fmul_rn(p0, x);      // :2014-2019 -> bit-exact
        case Map::Elu: return x > 0.0f ? x : __fmul_rn(p0, __fsub_rn(expf(x), 1.0f));   // :2118-2122
        case Map::Pow: {
            // :3342 x.powf(n).  The exponents the reference's own tests pin with assert_eq (2, 0.5, -1, 0, 1:
            // src/vector.rs test_pow_*) take the correctly rounded single operation a correctly rounded powf returns
            // for them; p0 is a launch parameter, so the branch is uniform.  Everything else: CUDA's powf (<= 4 ulp).
            if (p0 == 2.0f) return __fmul_rn(x, x);
            if (p0 == -1.0f) return __frcp_rn(x);
            if (p0 == 0.5f) return x > 0.0f ? sqrtf(x) : powf(x, 0.5f);   // pow(-0, .5) = +0 and pow(-inf, .5) = +inf, unlike sqrt
            if (p0 == 1.0f) return x;
            if (p0 == 0.0f) return 1.0f;
            return powf(x, p0);
        }
        case Map::Minimum: return fminf(x, y);                            // :4328 f32::min: the non-NaN operand
        case Map::Maximum: return fmaxf(x, y);                            // :4364
        case Map::Copysign: return copysignf(x, y);                       // :4292
        case Map::Affine: return __fmul_rn(__fsub_rn(x, p0), p1);         // (x - mean) * inv_std, :1195-1200, :1265-1270
    }
    return x;
}

template <Map OP>
__device__ __forceinline__ float4 apply4(const float4& x, const float4& y, const float4& z, float p0, float p1) {
    return make_float4(apply<OP>(x.x, y.x, z.x, p0, p1), apply<OP>(x.y, y.y, z.y, p0, p1),
                       apply<OP>(x.z, y.z, z.z, p0, p1), apply<OP>(x.w, y.w, z.w, p0, p1));
}

template <Map OP, bool VEC>
__global__ void __launch_bounds__(kThreads)
map_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c,
           float* __restrict__ out, size_t n, float p0, float p1) {
    constexpr int AR = map_arity(OP);
    if (VEC) {
        const size_t nvec = n >> 2;
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        const float4* c4 = reinterpret_cast<const float4*>(c);
        float4* o4 = reinterpret_cast<float4*>(out);
        const size_t base = (size_t)blockIdx.x * (kThreads * kUnroll) + threadIdx.x;
        if (base + (kUnroll - 1) * kThreads < nvec) {   // full tile (all but the last CTA)
            float4 x[kUnroll], y[kUnroll], z[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(a4 + base + u * kThreads);
            if (AR >= 2) {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) y[u] = ld_stream(b4 + base + u * kThreads);
            }
            if (AR >= 3) {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) z[u] = ld_stream(c4 + base + u * kThreads);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u)
                st_stream(o4 + base + u * kThreads, apply4<OP>(x[u], AR >= 2 ? y[u] : x[u], AR >= 3 ? z[u] : x[u], p0, p1));
        } else {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const size_t v = base + u * kThreads;
                if (v < nvec) {
                    const float4 x = ld_stream(a4 + v);
                    const float4 y = AR >= 2 ? ld_stream(b4 + v) : x;
                    const float4 z = AR >= 3 ? ld_stream(c4 + v) : x;
                    st_stream(o4 + v, apply4<OP>(x, y, z, p0, p1));
                }
            }
        }
        if (blockIdx.x == 0) {   // scalar tail (n % 4 elements)
            const size_t i = (nvec << 2) + threadIdx.x;
            if (i < n) out[i] = apply<OP>(a[i], AR >= 2 ? b[i] : 0.f, AR >= 3 ? c[i] : 0.f, p0, p1);
        }
    } else {
        const size_t base = (size_t)blockIdx.x * (kThreads * kUnroll * 4) + threadIdx.x;
#pragma unroll
        for (int u = 0; u < kUnroll * 4; ++u) {
            const size_t i = base + (size_t)u * kThreads;
            if (i < n)
                out[i] = apply<OP>(ld_stream(a + i), AR >= 2 ? ld_stream(b + i) : 0.f, AR >= 3 ? ld_stream(c + i) : 0.f, p0, p1);
        }
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

template <Map OP>
static void launch_one(bool vec, unsigned grid, const float* a, const float* b, const float* c, float* out, size_t n,
                       float p0, float p1, cudaStream_t s) {
    if (vec) map_kernel<OP, true><<<grid, kThreads, 0, s>>>(a, b, c, out, n, p0, p1);
    else     map_kernel<OP, false><<<grid, kThreads, 0, s>>>(a, b, c, out, n, p0, p1);
}

int launch_map(Map op, const float* a, const float* b, const float* c3, float* out, size_t n, float p0, float p1,
               cudaStream_t s) {
    Context* cx = ctx();
    if (!cx) return TRN_GPU_ERROR;
    if (n == 0) return TRN_OK;
    const int ar = map_arity(op);
    const bool vec = aligned16(a) && aligned16(out) && (ar < 2 || aligned16(b)) && (ar < 3 || aligned16(c3));
    // one CTA per 8 KiB tile; the scalar path covers the same 2048 elements per CTA
    const size_t per_cta = (size_t)kThreads * kUnroll * 4;
    const size_t tiles = (n + per_cta - 1) / per_cta;
    if (tiles > 0x7FFFFFFFull) return fail(TRN_INVALID_INPUT, "vector of %zu elements exceeds the launch grid", n);
    const unsigned grid = (unsigned)tiles;
#define CASE(OP) case Map::OP: launch_one<Map::OP>(vec, grid, a, b, c3, out, n, p0, p1, s); break
    switch (op) {
        CASE(Add); CASE(Sub); CASE(Mul); CASE(Div); CASE(Scale); CASE(Abs); CASE(Clamp); CASE(Lerp); CASE(Fma);
        CASE(Relu); CASE(Exp); CASE(Sigmoid); CASE(Gelu); CASE(Swish); CASE(Tanh); CASE(Sqrt); CASE(Recip);
        CASE(Ln); CASE(Log2); CASE(Log10); CASE(Sin); CASE(Cos); CASE(Tan); CASE(Floor); CASE(Ceil); CASE(Round);
        CASE(Neg); CASE(Signum); CASE(Trunc); CASE(Fract); CASE(Sinh); CASE(Cosh); CASE(Asin); CASE(Acos); CASE(Atan);
        CASE(Asinh); CASE(Acosh); CASE(Atanh); CASE(Hardswish); CASE(Mish); CASE(Selu); CASE(LeakyRelu); CASE(Elu);
        CASE(Pow); CASE(Minimum); CASE(Maximum); CASE(Copysign); CASE(Affine);
    }
#undef CASE
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

}  // namespace trn
