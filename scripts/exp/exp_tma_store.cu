// Experiment: per-SM throughput of cp.async.bulk.tensor stores for different box shapes / swizzles.
// Measured on B200 (store-only, 4 GiB): 32x32 f32 SWIZZLE_128B boxes 5.07 TB/s regardless of 1/2/4 staging buffers,
// 64x32 4.73, 128x32 4.60, un-swizzled 32x64 / 32x128 boxes 5.8 TB/s.  The GEMM epilogue's TMA stores therefore cost
// >= 0.85 ms per 4.3 GB of C by themselves and share the TMA unit with the operand loads.
// Each of 148 CTAs (8 warps) repeatedly stores its staging blocks to distinct tiles of a big [rows][cols] f32 tensor.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("err %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} }while(0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// each warp owns a staging block of BOX_R x BOX_C floats and stores it `iters` times to successive tiles
template <int BOX_R, int BOX_C, int NBUF>
__global__ void __launch_bounds__(256, 1) k_store(const __grid_constant__ CUtensorMap map, int tiles_per_row, int iters) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kBytes = BOX_R * BOX_C * 4;
    const uint32_t base = ((smem_u32(smem) + 1023u) & ~1023u) + warp * NBUF * kBytes;
    // fill staging once
    for (int i = lane; i < NBUF * kBytes / 4; i += 32) asm volatile("st.shared.f32 [%0], %1;" :: "r"(base + i * 4), "f"(1.0f));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0) {
        int buf = 0;
        for (int it = 0; it < iters; ++it) {
            const int tile = (blockIdx.x * 8 + warp) * iters + it;
            const int tr = tile / tiles_per_row, tc = tile % tiles_per_row;
            asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(NBUF - 1) : "memory");
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         :: "l"(&map), "r"(base + buf * kBytes), "r"(tc * BOX_C), "r"(tr * BOX_R) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            buf = (buf + 1) % NBUF;
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

template <int BOX_R, int BOX_C, int NBUF>
void run(EncodeTiledFn enc, float* d, size_t rows, size_t cols, CUtensorMapSwizzle sw, const char* name) {
    CUtensorMap map;
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * sizeof(float)};
    cuuint32_t box[2] = {BOX_C, BOX_R};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("%-40s encode failed %d\n", name, (int)r); return; }
    const int smem_bytes = 8 * NBUF * BOX_R * BOX_C * 4 + 1024;
    CK(cudaFuncSetAttribute(k_store<BOX_R, BOX_C, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    const int tiles_per_row = (int)(cols / BOX_C);
    const size_t total_tiles = (rows / BOX_R) * tiles_per_row;
    const int iters = (int)(total_tiles / (148 * 8));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<float> ts;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        k_store<BOX_R, BOX_C, NBUF><<<148, 256, smem_bytes>>>(map, tiles_per_row, iters);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep) ts.push_back(ms);
    }
    std::sort(ts.begin(), ts.end());
    const double bytes = (double)iters * 148 * 8 * BOX_R * BOX_C * 4;
    printf("%-40s %.3f ms  %.0f GB/s  (%.1f B/clk/SM @1.7GHz)\n", name, ts[ts.size() / 2], bytes / ts[ts.size() / 2] / 1e6,
           bytes / ts[ts.size() / 2] / 1e6 * 1e9 / 148 / 1.7e9 / 1e0 / 1.0);
}

int main() {
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)p;
    const size_t rows = 65536, cols = 16384;   // 4 GiB
    float* d; CK(cudaMalloc(&d, rows * cols * 4));
    run<32, 32, 1>(enc, d, rows, cols, CU_TENSOR_MAP_SWIZZLE_128B, "32x32 sw128 nbuf1 (current epilogue)");
    run<32, 32, 2>(enc, d, rows, cols, CU_TENSOR_MAP_SWIZZLE_128B, "32x32 sw128 nbuf2");
    run<32, 32, 4>(enc, d, rows, cols, CU_TENSOR_MAP_SWIZZLE_128B, "32x32 sw128 nbuf4");
    run<64, 32, 2>(enc, d, rows, cols, CU_TENSOR_MAP_SWIZZLE_128B, "64x32 sw128 nbuf2");
    run<128, 32, 1>(enc, d, rows, cols, CU_TENSOR_MAP_SWIZZLE_128B, "128x32 sw128 nbuf1");
    run<32, 64, 2>(enc, d, rows, cols, CU_TENSOR_MAP_SWIZZLE_NONE, "32x64 none nbuf2");
    run<32, 128, 1>(enc, d, rows, cols, CU_TENSOR_MAP_SWIZZLE_NONE, "32x128 none nbuf1");
    run<32, 128, 2>(enc, d, rows, cols, CU_TENSOR_MAP_SWIZZLE_NONE, "32x128 none nbuf2");
    // (16x256 / 8x256 boxes are rejected at launch: > 227 KiB of staging or an invalid box for this tensor)
    run<32, 32, 2>(enc, d, rows, cols, CU_TENSOR_MAP_SWIZZLE_NONE, "32x32 none nbuf2");
    return 0;
}
