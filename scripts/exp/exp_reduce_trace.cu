// Experiment: where do the ~8 us of fixed cost of the persistent sum kernel go?  Per-CTA timestamps (globaltimer) of kernel
// entry, end of the streaming loop and end of the block fold, plus the last block's fold, for one launch in a back-to-back
// series.  nvcc -arch=sm_100a -O3 -o /tmp/rt scripts/exp/exp_reduce_trace.cu && /tmp/rt [log2 n] [ctas per SM] [unroll variant]
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)
constexpr int kThreads = 256, kUnroll = 4, kTileVec = kThreads * kUnroll;
__device__ __forceinline__ float4 ld_stream(const float4* p){ float4 v; asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];":"=f"(v.x),"=f"(v.y),"=f"(v.z),"=f"(v.w):"l"(p)); return v;}
__device__ __forceinline__ unsigned long long gns(){ unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;":"=l"(t)); return t; }
__device__ __forceinline__ float warp_sum(float v){ for(int o=16;o>0;o>>=1) v+=__shfl_xor_sync(0xffffffffu,v,o); return v; }
__device__ __forceinline__ float block_sum(float v){ __shared__ float s_w[8]; v=warp_sum(v); if((threadIdx.x&31)==0) s_w[threadIdx.x>>5]=v; __syncthreads(); float r=0.f; if(threadIdx.x<32){ r=threadIdx.x<8?s_w[threadIdx.x]:0.f; r=warp_sum(r);} __syncthreads(); return r; }

// MODE 0: the library's loop (4 loads, consume, next).  MODE 1: software pipelined: next tile's loads issued before this tile is consumed.
// MODE 2: DYNAMIC chunks of `chunk` tiles claimed with an atomic counter (claimed one ahead), one partial per CHUNK (so the
// result does not depend on which CTA took which chunk); pdl: 1 = griddepcontrol trigger at kernel start, 2 = after the loop.
template <int MODE>
__global__ void __launch_bounds__(kThreads) sum_kernel(const float* __restrict__ a, size_t n, float* partial, unsigned* ticket, float* out, unsigned long long* ts,
                                                       int pdl, unsigned chunk, unsigned* claim) {
    if (pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (pdl == 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const unsigned long long t0 = gns();
    if (MODE == 3) {
        // order-free reductions (max / min / arg): chunks are claimed, the accumulators simply carry on — no per-chunk fold.
        // The next claim is fetched one chunk ahead; the only block-wide step per chunk is one barrier that publishes it.
        __shared__ unsigned s_nx[2];
        const size_t nvec = n >> 2;
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const unsigned ntiles = (unsigned)(nvec / kTileVec), nchunks = (ntiles + chunk - 1) / chunk;
        float acc[kUnroll] = {0.f,0.f,0.f,0.f};
        unsigned c = blockIdx.x, it = 0;
        if (threadIdx.x == 0) s_nx[0] = atomicAdd(claim, 1u) + gridDim.x;
        while (c < nchunks) {
            const unsigned t_end = min(ntiles, (c + 1) * chunk);
            for (unsigned t = c * chunk; t < t_end; ++t) {
                const size_t base = (size_t)t * kTileVec + threadIdx.x;
                float4 x[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(a4 + base + u * kThreads);
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) acc[u] = fmaxf(acc[u], fmaxf(fmaxf(x[u].x, x[u].y), fmaxf(x[u].z, x[u].w)));
            }
            __syncthreads();
            c = s_nx[it & 1];
            ++it;
            if (threadIdx.x == 0 && c < nchunks) s_nx[it & 1] = atomicAdd(claim, 1u) + gridDim.x;
        }
        const unsigned long long t1 = gns();
        float r = block_sum(fmaxf(fmaxf(acc[0], acc[1]), fmaxf(acc[2], acc[3])));
        if (threadIdx.x == 0) partial[blockIdx.x] = r;
        __shared__ bool s_last3;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) { unsigned t = atomicAdd(ticket, 1u); s_last3 = (t == gridDim.x - 1); if (s_last3) { *ticket = 0; *claim = 0; } }
        __syncthreads();
        const unsigned long long t2 = gns();
        if (threadIdx.x == 0) { ts[3 * blockIdx.x] = t0; ts[3 * blockIdx.x + 1] = t1; ts[3 * blockIdx.x + 2] = t2; }
        if (s_last3) {
            __threadfence();
            float r4 = 0.f;
            for (unsigned i = threadIdx.x; i < gridDim.x; i += kThreads) r4 = fmaxf(r4, __ldcg(partial + i));
            float rr = block_sum(r4);
            if (threadIdx.x == 0) { *out = rr; ts[3 * gridDim.x] = gns(); }
        }
        return;
    }
    if (MODE == 2) {
        __shared__ unsigned s_next[2];
        const size_t nvec = n >> 2;
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const unsigned ntiles = (unsigned)(nvec / kTileVec), nchunks = (ntiles + chunk - 1) / chunk;
        unsigned c = blockIdx.x;            // first chunk: static
        if (threadIdx.x == 0) s_next[0] = atomicAdd(claim, 1u) + gridDim.x;
        unsigned it = 0;
        while (c < nchunks) {
            float acc[kUnroll] = {0.f,0.f,0.f,0.f};
            const unsigned t_end = min(ntiles, (c + 1) * chunk);
            for (unsigned t = c * chunk; t < t_end; ++t) {
                const size_t base = (size_t)t * kTileVec + threadIdx.x;
                float4 x[kUnroll];
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(a4 + base + u * kThreads);
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) acc[u] += (x[u].x + x[u].y) + (x[u].z + x[u].w);
            }
            float r = block_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));   // two __syncthreads inside
            if (threadIdx.x == 0) partial[c] = r;
            c = s_next[it & 1];             // written before the first barrier of block_sum above
            ++it;
            if (threadIdx.x == 0 && c < nchunks) s_next[it & 1] = atomicAdd(claim, 1u) + gridDim.x;
        }
        if (pdl == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        const unsigned long long t1 = gns();
        __shared__ bool s_last2;
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) { unsigned t = atomicAdd(ticket, 1u); s_last2 = (t == gridDim.x - 1); if (s_last2) { *ticket = 0; *claim = 0; } }
        __syncthreads();
        const unsigned long long t2 = gns();
        if (threadIdx.x == 0) { ts[3 * blockIdx.x] = t0; ts[3 * blockIdx.x + 1] = t1; ts[3 * blockIdx.x + 2] = t2; }
        if (s_last2) {
            __threadfence();
            float r4[4] = {0.f,0.f,0.f,0.f};
            for (unsigned i = threadIdx.x; i < nchunks; i += 4 * kThreads) {
                float q[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) q[u] = i + u * kThreads < nchunks ? __ldcg(partial + i + u * kThreads) : 0.f;
#pragma unroll
                for (int u = 0; u < 4; ++u) r4[u] += q[u];
            }
            float rr = block_sum((r4[0] + r4[1]) + (r4[2] + r4[3]));
            if (threadIdx.x == 0) { *out = rr; ts[3 * gridDim.x] = gns(); }
        }
        return;
    }
    float acc[kUnroll] = {0.f,0.f,0.f,0.f};
    const size_t nvec = n >> 2;
    const float4* a4 = reinterpret_cast<const float4*>(a);
    const size_t full_tiles = (nvec / kTileVec) / gridDim.x * gridDim.x;
    if (MODE == 0) {
        for (size_t t = blockIdx.x; t < full_tiles; t += gridDim.x) {
            const size_t base = t * kTileVec + threadIdx.x;
            float4 x[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(a4 + base + u * kThreads);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) acc[u] += (x[u].x + x[u].y) + (x[u].z + x[u].w);
        }
    } else {
        size_t t = blockIdx.x;
        float4 x[kUnroll];
        if (t < full_tiles) {
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(a4 + t * kTileVec + threadIdx.x + u * kThreads);
        }
        for (; t < full_tiles; t += gridDim.x) {
            float4 y[kUnroll];
            const size_t tn = t + gridDim.x;
            if (tn < full_tiles) {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) y[u] = ld_stream(a4 + tn * kTileVec + threadIdx.x + u * kThreads);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) acc[u] += (x[u].x + x[u].y) + (x[u].z + x[u].w);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) x[u] = y[u];
        }
    }
    const size_t stride = (size_t)gridDim.x * kThreads;
    size_t v = full_tiles * kTileVec + (size_t)blockIdx.x * kThreads + threadIdx.x;
    for (; v + (kUnroll - 1) * stride < nvec; v += kUnroll * stride) {
        float4 x[kUnroll];
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(a4 + v + u * stride);
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) acc[u] += (x[u].x + x[u].y) + (x[u].z + x[u].w);
    }
    for (; v < nvec; v += stride) { float4 x = ld_stream(a4 + v); acc[0] += (x.x + x.y) + (x.z + x.w); }
    if (pdl == 2) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const unsigned long long t1 = gns();
    float r = block_sum((acc[0] + acc[1]) + (acc[2] + acc[3]));
    if (threadIdx.x == 0) partial[blockIdx.x] = r;
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) { unsigned t = atomicAdd(ticket, 1u); s_last = (t == gridDim.x - 1); if (s_last) *ticket = 0; }
    __syncthreads();
    const unsigned long long t2 = gns();
    if (threadIdx.x == 0) { ts[3 * blockIdx.x] = t0; ts[3 * blockIdx.x + 1] = t1; ts[3 * blockIdx.x + 2] = t2; }
    if (s_last) {
        __threadfence();
        float r4[4] = {0.f,0.f,0.f,0.f};
        for (unsigned i = threadIdx.x; i < gridDim.x; i += 4 * kThreads) {
            float q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = i + u * kThreads < gridDim.x ? __ldcg(partial + i + u * kThreads) : 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u) r4[u] += q[u];
        }
        float rr = block_sum((r4[0] + r4[1]) + (r4[2] + r4[3]));
        if (threadIdx.x == 0) { *out = rr; ts[3 * gridDim.x] = gns(); }
    }
}

int main(int argc, char** argv) {
    const int lg = argc > 1 ? atoi(argv[1]) : 27, per_sm = argc > 2 ? atoi(argv[2]) : 4, mode = argc > 3 ? atoi(argv[3]) : 0;
    const int pdl = argc > 4 ? atoi(argv[4]) : 0; const unsigned chunk = argc > 5 ? atoi(argv[5]) : 16;
    const size_t n = (size_t)1 << lg;
    float* a; CK(cudaMalloc(&a, n * 4)); CK(cudaMemset(a, 0, n * 4));
    const int grid = 148 * per_sm;
    float *partial, *out; unsigned* ticket; unsigned long long* ts;
    CK(cudaMalloc(&partial, (grid + (n >> 12) / chunk + 16) * 4)); unsigned* claim; CK(cudaMalloc(&claim, 4)); CK(cudaMemset(claim, 0, 4)); CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&ticket, 4)); CK(cudaMemset(ticket, 0, 4));
    CK(cudaMalloc(&ts, (3 * grid + 1) * 8));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto launch = [&] {
        cudaLaunchConfig_t cfg = {}; cfg.gridDim = grid; cfg.blockDim = kThreads;
        cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
        cfg.attrs = at; cfg.numAttrs = 1;
        if (mode == 0) CK(cudaLaunchKernelEx(&cfg, sum_kernel<0>, (const float*)a, n, partial, ticket, out, ts, pdl, chunk, claim));
        else if (mode == 1) CK(cudaLaunchKernelEx(&cfg, sum_kernel<1>, (const float*)a, n, partial, ticket, out, ts, pdl, chunk, claim));
        else if (mode == 2) CK(cudaLaunchKernelEx(&cfg, sum_kernel<2>, (const float*)a, n, partial, ticket, out, ts, pdl, chunk, claim));
        else CK(cudaLaunchKernelEx(&cfg, sum_kernel<3>, (const float*)a, n, partial, ticket, out, ts, pdl, chunk, claim));
    };
    for (int i = 0; i < 20; ++i) launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0)); for (int i = 0; i < 40; ++i) launch(); CK(cudaEventRecord(e1)); CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<unsigned long long> h(3 * grid + 1);
    CK(cudaMemcpy(h.data(), ts, h.size() * 8, cudaMemcpyDeviceToHost));
    std::vector<double> s(grid), l(grid), d(grid);
    unsigned long long t0 = ~0ull; for (int i = 0; i < grid; ++i) t0 = std::min(t0, h[3 * i]);
    for (int i = 0; i < grid; ++i) { s[i] = (h[3 * i] - t0) * 1e-3; l[i] = (h[3 * i + 1] - t0) * 1e-3; d[i] = (h[3 * i + 2] - t0) * 1e-3; }
    std::sort(s.begin(), s.end()); std::sort(l.begin(), l.end()); std::sort(d.begin(), d.end());
    auto q = [&](std::vector<double>& v, double f) { return v[(size_t)(f * (v.size() - 1))]; };
    printf("n=2^%d grid=%d mode=%d pdl=%d chunk=%u: %.1f us per launch (back to back), ideal at 7.24 TB/s %.1f us\n", lg, grid, mode, pdl, chunk, ms * 1e3 / 40, 4.0 * n / 7.24e6);
    printf("  CTA start      us after the first: min %.2f  median %.2f  p90 %.2f  max %.2f\n", q(s, 0), q(s, .5), q(s, .9), q(s, 1));
    printf("  stream loop end: min %.2f  p10 %.2f  median %.2f  p90 %.2f  max %.2f\n", q(l, 0), q(l, .1), q(l, .5), q(l, .9), q(l, 1));
    printf("  block fold+ticket done: min %.2f median %.2f max %.2f;  last block wrote the result at %.2f\n", q(d, 0), q(d, .5), q(d, 1), (h[3 * grid] - t0) * 1e-3);
    return 0;
}
