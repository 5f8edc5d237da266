// map.cu — elementwise kernels: add, mul (bit-exact), sigmoid, gelu (scalar-backend definitions).
//
// Replaces Avx2Backend::{add,mul,sigmoid,gelu} (src/backends/avx2.rs:32-117, :875-1047).  add/mul
// are one IEEE operation per element, so they match the reference bit for bit.  sigmoid/gelu follow
// the scalar backend's definitions (src/backends/scalar.rs:313-340: libm expf with +-50 cut-offs;
// tanh-form GELU with libm tanhf) using CUDA's accurate expf/tanhf and IEEE division — no
// fast-math intrinsics on the parity path.
//
// HBM-bound streaming: 128-bit loads/stores, 4 vectors in flight per thread per input, grid-stride
// over a persistent grid.  Algorithmic bytes per element: add/mul 12 B, sigmoid/gelu 8 B.
#include "common.cuh"

namespace trn {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

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
        constexpr int tile = kThreads * kUnroll;
        const size_t full_tiles = nvec / tile;
        for (size_t t = blockIdx.x; t < full_tiles; t += gridDim.x) {
            const size_t base = t * tile + threadIdx.x;
            float4 x[kUnroll], y[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(a4 + base + u * kThreads);
            if (BIN) {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) y[u] = ld_stream(b4 + base + u * kThreads);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) st_stream(o4 + base + u * kThreads, apply4<OP>(x[u], BIN ? y[u] : x[u]));
        }
        const size_t tail0 = full_tiles * tile;
        for (size_t v = tail0 + (size_t)blockIdx.x * kThreads + threadIdx.x; v < nvec; v += (size_t)gridDim.x * kThreads) {
            float4 x = ld_stream(a4 + v);
            float4 y = BIN ? ld_stream(b4 + v) : x;
            st_stream(o4 + v, apply4<OP>(x, y));
        }
        if (blockIdx.x == 0) {
            size_t i = (nvec << 2) + threadIdx.x;
            if (i < n) out[i] = apply<OP>(a[i], BIN ? b[i] : 0.f);
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads)
            out[i] = apply<OP>(ld_stream(a + i), BIN ? ld_stream(b + i) : 0.f);
    }
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int launch_map(Map op, const float* a, const float* b, float* out, size_t n, cudaStream_t s) {
    Context* c = ctx();
    if (!c) return TRN_GPU_ERROR;
    if (n == 0) return TRN_OK;
    const bool bin = op == Map::Add || op == Map::Mul;
    const bool vec = aligned16(a) && aligned16(out) && (!bin || aligned16(b));
    size_t tiles = (n / 4 + kThreads * kUnroll - 1) / (kThreads * kUnroll);
    // exactly one resident wave: grid = SMs x (CTAs the kernel really fits per SM), so there is no tail wave
#define LAUNCH(OP)                                                                    \
    do {                                                                              \
        static int per_sm = 0;                                                        \
        if (!per_sm) {                                                                \
            TRN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, map_kernel<OP, true>, kThreads, 0)); \
            if (per_sm < 1) per_sm = 1;                                               \
        }                                                                             \
        size_t cap = (size_t)c->sm_count * per_sm;                                    \
        int grid = (int)(tiles < cap ? (tiles ? tiles : 1) : cap);                    \
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
