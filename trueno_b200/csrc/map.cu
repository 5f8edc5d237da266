// map.cu — elementwise kernels: add, mul (bit-exact), sigmoid, gelu (scalar-backend definitions).
//
// Replaces Avx2Backend::{add,mul,sigmoid,gelu} (src/backends/avx2.rs:32-117, :875-1047).  add/mul
// are one IEEE operation per element, so they match the reference bit for bit.  sigmoid/gelu follow
// the scalar backend's definitions (src/backends/scalar.rs:313-340: libm expf with +-50 cut-offs;
// tanh-form GELU with libm tanhf) using CUDA's accurate expf/tanhf and IEEE division — no
// fast-math intrinsics on the parity path.
//
// HBM-bound streaming: 128-bit loads/stores, ONE 8 KiB tile per CTA on a flat (non-persistent) grid.
// Measured on B200 (scripts/exp/exp_map.cu, 131 M elements): flat grids reach 6.6-6.9 TB/s where a
// persistent one-wave grid-stride loop of the same body reaches 5.0-6.5 TB/s — the block scheduler
// keeps every SM's load queue full while persistent CTAs drift into lock-step load/compute phases.
// Algorithmic bytes per element: add/mul 12 B, sigmoid/gelu 8 B.
#include "common.cuh"

namespace trn {

constexpr int kThreads = 256;
constexpr int kUnroll = 2;   // float4 per thread per input: 8 KiB tiles

template <int OP>
__device__ __forceinline__ float apply(float x, float y) {
    if (OP == 0) return x + y;
    if (OP == 1) return x * y;
    if (OP == 2) {
        // src/backends/scalar.rs:313-324 (cut-offs applied as selects: no divergence)
        const float s = 1.0f / (1.0f + expf(-x));
        return x < -50.0f ? 0.0f : (x > 50.0f ? 1.0f : s);
    }
    // src/backends/scalar.rs:330-340: 0.5*x*(1 + tanh(u)), u = k*(x + c*x^3) with the reference's
    // operation order for u.  Evaluated through the exact identity 0.5*(1 + tanh(u)) = 1/(1 + e^(-2u)):
    // branch-free (tanhf is two divergent paths), ~half the instructions, and WITHOUT the 1 + tanh
    // cancellation the reference's form has for x << 0 — so the result is at least as close to the true
    // value as the reference's own (parity tolerance: 4 ulp + 4*2^-24*|x|, the reference's cancellation term).
    const float x3 = __fmul_rn(__fmul_rn(x, x), x);
    const float u = __fmul_rn(0.7978846f, __fadd_rn(x, __fmul_rn(0.044715f, x3)));
    return x / (1.0f + expf(-2.0f * u));
}

template <int OP>
__device__ __forceinline__ float4 apply4(const float4& x, const float4& y) {
    return make_float4(apply<OP>(x.x, y.x), apply<OP>(x.y, y.y), apply<OP>(x.z, y.z), apply<OP>(x.w, y.w));
}

template <int OP, bool VEC>
__global__ void __launch_bounds__(kThreads)
map_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, size_t n) {
    constexpr bool BIN = OP < 2;
    if (VEC) {
        const size_t nvec = n >> 2;
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        float4* o4 = reinterpret_cast<float4*>(out);
        const size_t base = (size_t)blockIdx.x * (kThreads * kUnroll) + threadIdx.x;
        if (base + (kUnroll - 1) * kThreads < nvec) {   // full tile (all but the last CTA)
            float4 x[kUnroll], y[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(a4 + base + u * kThreads);
            if (BIN) {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) y[u] = ld_stream(b4 + base + u * kThreads);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) st_stream(o4 + base + u * kThreads, apply4<OP>(x[u], BIN ? y[u] : x[u]));
        } else {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const size_t v = base + u * kThreads;
                if (v < nvec) {
                    const float4 x = ld_stream(a4 + v);
                    st_stream(o4 + v, apply4<OP>(x, BIN ? ld_stream(b4 + v) : x));
                }
            }
        }
        if (blockIdx.x == 0) {   // scalar tail (n % 4 elements)
            const size_t i = (nvec << 2) + threadIdx.x;
            if (i < n) out[i] = apply<OP>(a[i], BIN ? b[i] : 0.f);
        }
    } else {
        const size_t base = (size_t)blockIdx.x * (kThreads * kUnroll * 4) + threadIdx.x;
#pragma unroll
        for (int u = 0; u < kUnroll * 4; ++u) {
            const size_t i = base + (size_t)u * kThreads;
            if (i < n) out[i] = apply<OP>(ld_stream(a + i), BIN ? ld_stream(b + i) : 0.f);
        }
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int launch_map(Map op, const float* a, const float* b, float* out, size_t n, cudaStream_t s) {
    Context* c = ctx();
    if (!c) return TRN_GPU_ERROR;
    if (n == 0) return TRN_OK;
    const bool bin = op == Map::Add || op == Map::Mul;
    const bool vec = aligned16(a) && aligned16(out) && (!bin || aligned16(b));
    // one CTA per 8 KiB tile; the scalar path covers the same 2048 elements per CTA
    const size_t per_cta = (size_t)kThreads * kUnroll * 4;
    const size_t tiles = (n + per_cta - 1) / per_cta;
    if (tiles > 0x7FFFFFFFull) return fail(TRN_INVALID_INPUT, "vector of %zu elements exceeds the launch grid", n);
    const unsigned grid = (unsigned)tiles;
#define LAUNCH(OP)                                                                    \
    do {                                                                              \
        if (vec) map_kernel<OP, true><<<grid, kThreads, 0, s>>>(a, b, out, n);         \
        else     map_kernel<OP, false><<<grid, kThreads, 0, s>>>(a, b, out, n);        \
    } while (0)
    switch (op) {
        case Map::Add:     LAUNCH(0); break;
        case Map::Mul:     LAUNCH(1); break;
        case Map::Sigmoid: LAUNCH(2); break;
        case Map::Gelu:    LAUNCH(3); break;
    }
#undef LAUNCH
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

}  // namespace trn
