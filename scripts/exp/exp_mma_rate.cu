// Probe: cycles per tcgen05.mma kind::tf32 (M = 128, K = 8) issued back to back on shared-memory-resident operands,
// by N, by where A comes from (shared memory vs tensor memory) and by swizzle width.  No TMA, no barriers in the loop.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../trueno_b200/csrc/tcgen05.cuh"
using namespace trn;
using namespace trn::tc;

template <int SBK>
__global__ void __launch_bounds__(128, 1) probe(int n, int a_from_tmem, int iters, long long* out) {
    extern __shared__ uint8_t raw[];
    const uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(raw + (base - smem_u32(raw)))[i] = 0;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&slot), 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_tf32(128, n);
        // 4 stages x {A_hi, B_hi, A_lo, B_lo}, each tile 256 rows x SBK floats (enough for N = 256)
        constexpr uint32_t tile = 256 * SBK * 4;
        if (a_from_tmem == 2) {
            // bare issue rate: descriptors precomputed, nothing but MMAs in the loop
            const uint64_t d0 = make_desc_k<SBK>(base), d1 = make_desc_k<SBK>(base + tile), d2 = make_desc_k<SBK>(base + 2 * tile),
                           d3 = make_desc_k<SBK>(base + 3 * tile);
            const uint64_t st = (4 * tile) >> 4;   // second stage
            long long t0 = clock64();
#pragma unroll 1
            for (int it = 0; it < iters; it += 2) {
                umma_tf32(tm, d2, d1, idesc, 1u);
                umma_tf32(tm, d0, d3, idesc, 1u);
                umma_tf32(tm, d0, d1, idesc, 1u);
                umma_tf32(tm, d2 + 2, d1 + 2, idesc, 1u);
                umma_tf32(tm, d0 + 2, d3 + 2, idesc, 1u);
                umma_tf32(tm, d0 + 2, d1 + 2, idesc, 1u);
                umma_tf32(tm, d2 + st, d1 + st, idesc, 1u);
                umma_tf32(tm, d0 + st, d3 + st, idesc, 1u);
                umma_tf32(tm, d0 + st, d1 + st, idesc, 1u);
                umma_tf32(tm, d2 + st + 2, d1 + st + 2, idesc, 1u);
                umma_tf32(tm, d0 + st + 2, d3 + st + 2, idesc, 1u);
                umma_tf32(tm, d0 + st + 2, d1 + st + 2, idesc, 1u);
            }
            umma_commit(smem_u32(&bar));
            mbar_wait(smem_u32(&bar), 0);
            long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = t1 - t0;
        } else {
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            const uint32_t sa = base + (it & 1) * 4 * tile;   // alternate between two stages
            const uint32_t a_hi = sa, b_hi = sa + tile, a_lo = sa + 2 * tile, b_lo = sa + 3 * tile;
#pragma unroll
            for (int k = 0; k < SBK / 8; ++k) {
                const uint32_t koff = k * 32;
                if (a_from_tmem) {
                    umma_tf32_ts(tm, tm + 256 + k * 8, make_desc_k<SBK>(b_hi + koff), idesc, 1u);
                    umma_tf32_ts(tm, tm + 384 + k * 8, make_desc_k<SBK>(b_lo + koff), idesc, 1u);
                    umma_tf32_ts(tm, tm + 384 + k * 8, make_desc_k<SBK>(b_hi + koff), idesc, 1u);
                } else {
                    umma_tf32(tm, make_desc_k<SBK>(a_lo + koff), make_desc_k<SBK>(b_hi + koff), idesc, 1u);
                    umma_tf32(tm, make_desc_k<SBK>(a_hi + koff), make_desc_k<SBK>(b_lo + koff), idesc, 1u);
                    umma_tf32(tm, make_desc_k<SBK>(a_hi + koff), make_desc_k<SBK>(b_hi + koff), idesc, 1u);
                }
            }
        }
        umma_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
    long long* d; cudaMalloc(&d, 8);
    const int smem = 161 * 1024 + 1024;
    cudaFuncSetAttribute(probe<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 2000;
    for (int sbk : {16})
        for (int from_tmem : {0, 1, 2})
            for (int n : {64, 128, 256}) {
                if (from_tmem && n == 256) { /* D occupies [0,256): A columns start at 256 */ }
                long long h = 0;
                for (int rep = 0; rep < 2; ++rep) {
                    if (sbk == 16) probe<16><<<148, 128, smem>>>(n, from_tmem, iters, d);
                    else           probe<32><<<148, 128, smem>>>(n, from_tmem, iters, d);
                    cudaError_t e = cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
                    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
                }
                const double mmas = (double)iters * (sbk / 8) * 3;
                printf("SBK %2d  A from %-4s  N %3d : %7.1f cycles per MMA  (math %3d)\n", sbk, from_tmem == 1 ? "TMEM" : from_tmem == 2 ? "bare" : "smem", n,
                       (double)h / mmas, n / 2);
            }
    return 0;
}
