// softmax.cu — row softmax / log_softmax with ONE read and ONE write of HBM per element.
//
// Replaces Vector::softmax / log_softmax (src/vector.rs:1516-1553, :1581-1623: max -> expf(x-max)
// -> sum -> divide, i.e. 3 reads + 2 writes per element on the CPU, 4 dispatches with a host round
// trip each in the wgpu path, src/backends/gpu/device.rs:956-971).
//
// Layout: a row lives entirely in REGISTERS, spread over a thread-block CLUSTER: CS CTAs x 256
// threads x VPT float4 (CS in {1,2,4,8}, VPT in {1,2,4,8}); config 5's 32 000-float rows use
// CS=4, VPT=8 (32 768 slots).  Row max and exp-sum are combined across the cluster through
// distributed shared memory in a fixed rank order, so every CTA computes bit-identical
// statistics and reruns are bit-identical.  Several clusters are resident per SM, so one row's
// load phase overlaps another row's exp/store phase.  Rows longer than 65 536 elements (or rows
// that are not 16-byte aligned) take the three-pass fallback kernel, which re-reads the row from
// L2.  Math: accurate expf / logf and IEEE division — no fast-math intrinsics.
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

}  // namespace trn
