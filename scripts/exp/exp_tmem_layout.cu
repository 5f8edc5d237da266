// Probe: register layout of tcgen05.ld.16x256b (written through 32x32b stores: lane = row, register = column).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(uint32_t* out) {
    __shared__ uint32_t slot;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(32u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = slot;
    if (warp == 0) {
        // row = lane (0..31), col = 0..15: value = row * 100 + col
        for (int c = 0; c < 16; ++c) {
            uint32_t v = lane * 100 + c;
            asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(base + c), "r"(v) : "memory");
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        uint32_t r[8];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(base) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) out[lane * 8 + i] = r[i];
        uint32_t q[8];
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7]) : "r"(base + (16u << 16)) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int i = 0; i < 8; ++i) out[256 + lane * 8 + i] = q[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(32u) : "memory");
}
int main() {
    uint32_t* d; cudaMalloc(&d, 512 * 4);
    probe<<<1, 32>>>(d);
    uint32_t h[512]; cudaError_t e = cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    printf("status %s\n", cudaGetErrorString(e));
    for (int half = 0; half < 2; ++half) for (int l = 0; l < 32; ++l) {
        printf("half %d lane %2d:", half, l);
        for (int i = 0; i < 8; ++i) printf(" %4u", h[half * 256 + l * 8 + i]);
        printf("\n");
    }
}
