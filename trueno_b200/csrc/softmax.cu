// softmax.cu — row softmax / log_softmax with ONE read and ONE write of HBM per element.
//
// Replaces Vector::softmax / log_softmax (src/vector.rs:1516-1553, :1581-1623: max -> expf(x-max)
// -> sum -> divide, i.e. 3 reads + 2 writes per element on the CPU, 4 dispatches with a host round
// trip each in the wgpu path, src/backends/gpu/device.rs:956-971).
//
// Three kernels, chosen by row length (cols % 4 == 0, 16-byte aligned rows):
//   * cols <= 8192: a row lives in the REGISTERS of one 256-thread CTA (VPT float4 per thread).
//   * 8192 < cols <= 32768 (config 5: 32 000): the TMA RING kernel.  One persistent CTA per SM;
//     a producer warp streams rows into a 14-slot x 16 KiB shared-memory ring with 1-D bulk copies
//     (cp.async.bulk + mbarrier complete_tx), 16 consumer warps pull each slot into registers
//     (a row = <= 8 slots = 16 float4 per thread), release the slot at once, and compute
//     max -> exp -> sum -> scale out of registers.  Because slots are released as soon as they are
//     in registers, the next row's bulk copies (up to 224 KiB in flight per SM) run underneath the
//     current row's exp and store phases — HBM reads never stop, with no register cost.
//   * cols <= 65536: a row spread over a thread-block CLUSTER (CS CTAs x 256 threads x VPT float4);
//     row max and exp-sum are combined through distributed shared memory in a fixed rank order.
// Rows longer than that (or rows that are not 16-byte aligned) take the three-pass fallback kernel,
// which re-reads the row from L2.  All reductions use fixed trees, so reruns are bit-identical.
// Math: accurate expf / logf; softmax scales by the correctly rounded reciprocal of the row sum
// (<= 1 ulp from the reference's e / sum) — no fast-math intrinsics.
//
// Algorithmic bytes per element: 8 B (4 read + 4 written).  HBM-bound.
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace trn {

constexpr int kThreads = 256;

__device__ __forceinline__ float block_max_256(float v, float* s_w) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = s_w[0];
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) r = fmaxf(r, s_w[w]);
    __syncthreads();
    return r;  // identical in every thread
}
__device__ __forceinline__ float block_sum_256(float v, float* s_w) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = s_w[0];
#pragma unroll
    for (int w = 1; w < kThreads / 32; ++w) r += s_w[w];
    __syncthreads();
    return r;
}

// One cluster per row (grid-strided over rows).  cols % 4 == 0 and 16-byte aligned rows.
template <int CS, int VPT, bool LOG>
__global__ void __launch_bounds__(kThreads)
softmax_rows_cluster_kernel(const float* __restrict__ in, float* __restrict__ out, size_t rows, size_t cols) {
    __shared__ float s_w[kThreads / 32];
    __shared__ float s_stat[2];  // [0] CTA max, [1] CTA exp-sum — read by cluster peers through DSMEM

    unsigned rank = 0;
    if (CS > 1) rank = cg::this_cluster().block_rank();
    const size_t cluster_id = blockIdx.x / CS;
    const size_t num_clusters = gridDim.x / CS;
    const unsigned nvec = (unsigned)(cols >> 2);

    for (size_t row = cluster_id; row < rows; row += num_clusters) {
        const float4* src = reinterpret_cast<const float4*>(in + row * cols);
        float4* dst = reinterpret_cast<float4*>(out + row * cols);

        // ---- load: VPT independent 128-bit loads per thread; chunk j of the row is split
        //      contiguously over the CS CTAs so every warp reads 512 contiguous bytes
        float4 x[VPT];
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const unsigned v = (j * CS + rank) * kThreads + threadIdx.x;
            x[j] = v < nvec ? ld_stream(src + v) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
        }

        // ---- row max
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < VPT; ++j) m = fmaxf(m, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
        m = block_max_256(m, s_w);
        if (CS > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            if (threadIdx.x == 0) s_stat[0] = m;
            cluster.sync();
            float r = *cluster.map_shared_rank(&s_stat[0], 0);
#pragma unroll
            for (int p = 1; p < CS; ++p) r = fmaxf(r, *cluster.map_shared_rank(&s_stat[0], p));
            m = r;
        }

        // ---- exponentials and their sum.  Padding slots hold -inf -> expf(-inf) = 0 exactly.
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            float4 e;
            e.x = expf(x[j].x - m); e.y = expf(x[j].y - m); e.z = expf(x[j].z - m); e.w = expf(x[j].w - m);
            part += (e.x + e.y) + (e.z + e.w);
            if (!LOG) x[j] = e;
        }
        float sum = block_sum_256(part, s_w);
        if (CS > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            if (threadIdx.x == 0) s_stat[1] = sum;
            cluster.sync();
            float r = *cluster.map_shared_rank(&s_stat[1], 0);
#pragma unroll
            for (int p = 1; p < CS; ++p) r += *cluster.map_shared_rank(&s_stat[1], p);
            sum = r;
        }

        // ---- normalise and store
        const float lse = LOG ? logf(sum) : 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const unsigned v = (j * CS + rank) * kThreads + threadIdx.x;
            if (v < nvec) {
                float4 y;
                if (LOG) {  // (x - max) - ln(sum), evaluated in that order (src/vector.rs:1617-1621)
                    y.x = (x[j].x - m) - lse; y.y = (x[j].y - m) - lse;
                    y.z = (x[j].z - m) - lse; y.w = (x[j].w - m) - lse;
                } else {
                    y.x = x[j].x / sum; y.y = x[j].y / sum; y.z = x[j].z / sum; y.w = x[j].w / sum;
                }
                st_stream(dst + v, y);
            }
        }
        // peers must be done reading this CTA's s_stat before the next row overwrites it
        if (CS > 1) cg::this_cluster().sync();
    }
}

// ---- TMA ring kernel -----------------------------------------------------------------------------------
namespace ring {
constexpr int kConsumers = 512;                 // 16 consumer warps
constexpr int kRingThreads = kConsumers + 32;   // + 1 producer warp
constexpr int kChunkVec = 2 * kConsumers;       // float4 per slot: two per consumer thread
constexpr uint32_t kChunkBytes = kChunkVec * 16;  // 16 KiB
constexpr int kMaxChunks = 8;                   // row <= 8 slots = 32 768 floats
constexpr int kSlots = 14;                      // 224 KiB ring
constexpr uint32_t kSmemBytes = kSlots * kChunkBytes + 2 * kSlots * 8 + 2 * 16 * 4 + 128;
}  // namespace ring

template <bool LOG>
__global__ void __launch_bounds__(ring::kRingThreads, 1)
softmax_rows_ring_kernel(const float* __restrict__ in, float* __restrict__ out, size_t rows, size_t cols) {
    using namespace ring;
    extern __shared__ uint8_t ring_smem_raw[];
    const uint32_t base = (smem_u32(ring_smem_raw) + 127u) & ~127u;
    uint8_t* gen = ring_smem_raw + (base - smem_u32(ring_smem_raw));
    const uint32_t bar_base = base + kSlots * kChunkBytes;
    auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (kSlots + s); };
    float* s_max = reinterpret_cast<float*>(gen + kSlots * kChunkBytes + 2 * kSlots * 8);
    float* s_sum = s_max + 16;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned nvec = (unsigned)(cols >> 2);
    const unsigned nchunks = (nvec + kChunkVec - 1) / kChunkVec;
    const uint32_t row_bytes = nvec * 16u;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < (uint32_t)kSlots; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), kConsumers / 32); }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == kConsumers / 32) {
        // ===================== producer: one elected lane streams rows into the ring =====================
        if (elect_one()) {
            uint32_t slot = 0, phase = 0;
            for (size_t row = blockIdx.x; row < rows; row += gridDim.x) {
                const char* src = reinterpret_cast<const char*>(in + row * cols);
                for (unsigned j = 0; j < nchunks; ++j) {
                    mbar_wait(empty_bar(slot), phase ^ 1);
                    const uint32_t off = j * kChunkBytes;
                    const uint32_t bytes = row_bytes - off < kChunkBytes ? row_bytes - off : kChunkBytes;
                    mbar_expect_tx(full_bar(slot), bytes);
                    bulk_load_1d(base + slot * kChunkBytes, src + off, bytes, full_bar(slot));
                    if (++slot == (uint32_t)kSlots) { slot = 0; phase ^= 1; }
                }
            }
        }
        return;
    }

    // ===================== consumers =====================
    const int t = threadIdx.x;
    uint32_t slot = 0, phase = 0;
    for (size_t row = blockIdx.x; row < rows; row += gridDim.x) {
        float4 x[2 * kMaxChunks];
        // ---- ring -> registers; each slot goes back to the producer as soon as it has been read
#pragma unroll
        for (int j = 0; j < kMaxChunks; ++j) {
            if ((unsigned)j < nchunks) {
                mbar_wait(full_bar(slot), phase);
                const uint32_t sb = base + slot * kChunkBytes;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const unsigned v = j * kChunkVec + h * kConsumers + t;
                    float4 r = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                    if (v < nvec)
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(sb + (h * kConsumers + t) * 16u));
                    x[2 * j + h] = r;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_bar(slot));
                if (++slot == (uint32_t)kSlots) { slot = 0; phase ^= 1; }
            } else {
                x[2 * j] = x[2 * j + 1] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            }
        }

        // ---- row max: warp tree, then a fixed-order fold of the 16 warp values in every thread
        float m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 2 * kMaxChunks; ++j) m = fmaxf(m, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
        m = warp_max(m);
        if (lane == 0) s_max[warp] = m;
        named_bar_sync(1, kConsumers);
        m = s_max[0];
#pragma unroll
        for (int w = 1; w < kConsumers / 32; ++w) m = fmaxf(m, s_max[w]);

        // ---- exponentials and their sum.  Padding slots hold -inf -> expf(-inf) = 0 exactly.
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < 2 * kMaxChunks; ++j) {
            float4 e;
            e.x = expf(x[j].x - m); e.y = expf(x[j].y - m); e.z = expf(x[j].z - m); e.w = expf(x[j].w - m);
            part += (e.x + e.y) + (e.z + e.w);
            if (!LOG) x[j] = e;
        }
        part = warp_sum(part);
        if (lane == 0) s_sum[warp] = part;
        named_bar_sync(1, kConsumers);
        float sum = s_sum[0];
#pragma unroll
        for (int w = 1; w < kConsumers / 32; ++w) sum += s_sum[w];
        // (s_max is rewritten only after the next row's first barrier... no: after THIS barrier every
        //  thread has read s_max; s_sum is rewritten after the next row's max barrier.)

        // ---- scale and store straight from registers (512 contiguous bytes per warp per store)
        const float lse = LOG ? logf(sum) : 0.f;
        const float inv = LOG ? 0.f : __frcp_rn(sum);
        float4* dst = reinterpret_cast<float4*>(out + row * cols);
#pragma unroll
        for (int j = 0; j < 2 * kMaxChunks; ++j) {
            const unsigned v = (j >> 1) * kChunkVec + (j & 1) * kConsumers + t;
            if (v < nvec) {
                float4 y;
                if (LOG) {  // (x - max) - ln(sum), evaluated in that order (src/vector.rs:1617-1621)
                    y.x = (x[j].x - m) - lse; y.y = (x[j].y - m) - lse;
                    y.z = (x[j].z - m) - lse; y.w = (x[j].w - m) - lse;
                } else {
                    y.x = x[j].x * inv; y.y = x[j].y * inv; y.z = x[j].z * inv; y.w = x[j].w * inv;
                }
                st_stream(dst + v, y);
            }
        }
    }
}

template <bool LOG>
static int launch_ring(const float* a, float* out, size_t rows, size_t cols, int sm_count, cudaStream_t s) {
    static const cudaError_t attr = cudaFuncSetAttribute(softmax_rows_ring_kernel<LOG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                         (int)ring::kSmemBytes);   // thread-safe, once
    TRN_CUDA(attr);
    const unsigned grid = (unsigned)(rows < (size_t)sm_count ? rows : (size_t)sm_count);
    softmax_rows_ring_kernel<LOG><<<grid, ring::kRingThreads, ring::kSmemBytes, s>>>(a, out, rows, cols);
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

// Fallback: any cols / alignment.  One CTA per row, three passes (max, exp-sum, write); passes 2
// and 3 hit L2 for rows that fit there.
template <bool LOG>
__global__ void __launch_bounds__(kThreads)
softmax_rows_generic_kernel(const float* __restrict__ in, float* __restrict__ out, size_t rows, size_t cols) {
    __shared__ float s_w[kThreads / 32];
    for (size_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const float* src = in + row * cols;
        float* dst = out + row * cols;
        float m = -INFINITY;
        for (size_t i = threadIdx.x; i < cols; i += kThreads) m = fmaxf(m, src[i]);
        m = block_max_256(m, s_w);
        // long per-thread chains: compensated (Kahan) summation keeps the row sum within ~1 ulp
        float part = 0.f, comp = 0.f;
        for (size_t i = threadIdx.x; i < cols; i += kThreads) {
            const float y = __fsub_rn(expf(src[i] - m), comp);
            const float t = __fadd_rn(part, y);
            comp = __fsub_rn(__fsub_rn(t, part), y);
            part = t;
        }
        const float sum = block_sum_256(part, s_w);
        const float lse = LOG ? logf(sum) : 0.f;
        for (size_t i = threadIdx.x; i < cols; i += kThreads)
            dst[i] = LOG ? (src[i] - m) - lse : expf(src[i] - m) / sum;
    }
}

template <int CS, int VPT, bool LOG>
static int launch_cluster(const float* a, float* out, size_t rows, size_t cols, int sm_count, cudaStream_t s) {
    // resident clusters: aim for ~2048 threads' worth of rows per SM, bounded by the row count
    size_t want = (size_t)sm_count * 8 / CS;
    size_t clusters = rows < want ? rows : want;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * CS));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TRN_CUDA(cudaLaunchKernelEx(&cfg, softmax_rows_cluster_kernel<CS, VPT, LOG>, a, out, rows, cols));
    count_launch();
    return TRN_OK;
}

template <bool LOG>
static int dispatch(const float* a, float* out, size_t rows, size_t cols, int sm_count, cudaStream_t s) {
    const bool vec_ok = (cols % 4 == 0) && ((reinterpret_cast<uintptr_t>(a) & 15u) == 0) &&
                        ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
    const size_t nvec = cols / 4;
    if (vec_ok && nvec <= (size_t)kThreads * 8 * 8) {
        // smallest (CS, VPT) whose CS*256*VPT float4 slots hold the row; prefer registers over cluster width
        if (nvec <= kThreads * 1)      return launch_cluster<1, 1, LOG>(a, out, rows, cols, sm_count, s);
        if (nvec <= kThreads * 2)      return launch_cluster<1, 2, LOG>(a, out, rows, cols, sm_count, s);
        if (nvec <= kThreads * 4)      return launch_cluster<1, 4, LOG>(a, out, rows, cols, sm_count, s);
        if (nvec <= kThreads * 8)      return launch_cluster<1, 8, LOG>(a, out, rows, cols, sm_count, s);
        if (nvec <= (size_t)ring::kChunkVec * ring::kMaxChunks) return launch_ring<LOG>(a, out, rows, cols, sm_count, s);
        if (nvec <= kThreads * 8 * 2)  return launch_cluster<2, 8, LOG>(a, out, rows, cols, sm_count, s);
        if (nvec <= kThreads * 8 * 4)  return launch_cluster<4, 8, LOG>(a, out, rows, cols, sm_count, s);
        return launch_cluster<8, 8, LOG>(a, out, rows, cols, sm_count, s);
    }
    size_t cap = (size_t)sm_count * 8;
    int grid = (int)(rows < cap ? rows : cap);
    softmax_rows_generic_kernel<LOG><<<grid, kThreads, 0, s>>>(a, out, rows, cols);
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

int launch_softmax_rows(int log_variant, const float* a, float* out, size_t rows, size_t cols, cudaStream_t s) {
    Context* c = ctx();
    if (!c) return TRN_GPU_ERROR;
    return log_variant ? dispatch<true>(a, out, rows, cols, c->sm_count, s)
                       : dispatch<false>(a, out, rows, cols, c->sm_count, s);
}

// ---- layer_norm rows --------------------------------------------------------------
// Vector::layer_norm (src/vector.rs:1316-1362): mean = sum / n; variance = sum((x - mean)^2) / n;
// inv_std = 1 / sqrt(variance + eps); y = gamma * (x - mean) * inv_std + beta, evaluated per element in exactly
// that (unfused) order.  `rows` independent vectors share gamma / beta.  One CTA per row: the first pass
// streams the row from HBM, the second and third hit L1/L2 (a row is <= a few hundred KiB), so HBM sees one
// read and one write per element.  Fixed trees -> deterministic.
__global__ void __launch_bounds__(kThreads)
layer_norm_rows_kernel(const float* __restrict__ in, const float* __restrict__ gamma, const float* __restrict__ beta,
                       float eps, float* __restrict__ out, size_t rows, size_t cols) {
    __shared__ float s_w[kThreads / 32];
    for (size_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const float* src = in + row * cols;
        float* dst = out + row * cols;
        float part = 0.f;
        for (size_t i = threadIdx.x; i < cols; i += kThreads) part += src[i];
        const float mean = block_sum_256(part, s_w) / (float)cols;
        part = 0.f;
        for (size_t i = threadIdx.x; i < cols; i += kThreads) {
            const float d = __fsub_rn(src[i], mean);
            part = __fadd_rn(part, __fmul_rn(d, d));
        }
        const float var = block_sum_256(part, s_w) / (float)cols;
        const float inv_std = 1.0f / sqrtf(var + eps);
        for (size_t i = threadIdx.x; i < cols; i += kThreads)
            dst[i] = __fadd_rn(__fmul_rn(__fmul_rn(gamma[i], __fsub_rn(src[i], mean)), inv_std), beta[i]);
    }
}

int launch_layer_norm_rows(const float* a, const float* gamma, const float* beta, float eps, float* out, size_t rows,
                           size_t cols, cudaStream_t s) {
    Context* c = ctx();
    if (!c) return TRN_GPU_ERROR;
    if (rows == 0 || cols == 0) return TRN_OK;
    const size_t cap = (size_t)c->sm_count * 8;
    layer_norm_rows_kernel<<<(unsigned)(rows < cap ? rows : cap), kThreads, 0, s>>>(a, gamma, beta, eps, out, rows, cols);
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

}  // namespace trn
