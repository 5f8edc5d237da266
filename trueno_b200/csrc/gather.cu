// gather.cu — Matrix::embedding_lookup (src/matrix.rs:2008-2041): out[r, :] = table[indices[r], :].
//
// The reference copies one row per index with copy_from_slice; here one launch gathers every row: a group of
// 2^k threads (as many as the row has 128-bit vectors, up to a whole 256-thread CTA) moves one output row with
// four independent 128-bit loads in flight per thread, several rows per CTA when rows are short.  Table reads go
// through the normal cached path (token ids repeat, hot rows stay in L2), output rows are streamed out.
// Pure byte movement -> bit-exact.  Algorithmic bytes: 8 B per output element (4 read + 4 written) + 8 B per index.
// HBM-bound.
#include "common.cuh"

namespace trn {

constexpr int kGatherThreads = 256;

template <bool VEC>
__global__ void __launch_bounds__(kGatherThreads)
gather_rows_kernel(const float* __restrict__ table, size_t rows, size_t cols, const uint64_t* __restrict__ idx, size_t n_idx,
                   float* __restrict__ out, unsigned tpr_shift) {
    const unsigned tpr = 1u << tpr_shift;                       // threads per output row
    const unsigned rows_per_cta = kGatherThreads >> tpr_shift;
    const size_t orow = (size_t)blockIdx.x * rows_per_cta + (threadIdx.x >> tpr_shift);
    if (orow >= n_idx) return;
    const unsigned t = threadIdx.x & (tpr - 1);
    const uint64_t srow = idx[orow];
    const bool ok = srow < rows;   // the host-slice entry point validates before the launch; resident callers get zero rows
    if (VEC) {
        const size_t nvec = cols >> 2;
        const float4* s = reinterpret_cast<const float4*>(table + (ok ? srow : 0) * cols);
        float4* d = reinterpret_cast<float4*>(out + orow * cols);
        for (size_t v = t; v < nvec; v += 4 * (size_t)tpr) {
            float4 x[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const size_t u = v + (size_t)j * tpr;
                x[j] = (ok && u < nvec) ? __ldg(s + u) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const size_t u = v + (size_t)j * tpr;
                if (u < nvec) st_stream(d + u, x[j]);
            }
        }
    } else {
        const float* s = table + (ok ? srow : 0) * cols;
        float* d = out + orow * cols;
        for (size_t c = t; c < cols; c += tpr) d[c] = ok ? __ldg(s + c) : 0.f;
    }
}

int launch_gather_rows(const float* table, size_t rows, size_t cols, const uint64_t* idx, size_t n_idx, float* out,
                       cudaStream_t s) {
    Context* c = ctx();
    if (!c) return TRN_GPU_ERROR;
    if (n_idx == 0 || cols == 0) return TRN_OK;
    const bool vec = (cols % 4 == 0) && (((reinterpret_cast<uintptr_t>(table) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0);
    const size_t units = vec ? cols / 4 : cols;                 // work items per row
    unsigned shift = 0;
    while (shift < 8 && ((size_t)1 << shift) * (vec ? 4 : 1) < units) ++shift;   // 4 vectors per thread before a second loop trip
    const size_t rows_per_cta = (size_t)kGatherThreads >> shift;
    const size_t grid = (n_idx + rows_per_cta - 1) / rows_per_cta;
    if (grid > 0x7FFFFFFFull) return fail(TRN_INVALID_INPUT, "%zu indices exceed the launch grid", n_idx);
    if (vec) gather_rows_kernel<true><<<(unsigned)grid, kGatherThreads, 0, s>>>(table, rows, cols, idx, n_idx, out, shift);
    else     gather_rows_kernel<false><<<(unsigned)grid, kGatherThreads, 0, s>>>(table, rows, cols, idx, n_idx, out, shift);
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

}  // namespace trn
