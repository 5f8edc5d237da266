// eigen.cu — SymmetricEigen on the device: parallel two-sided Jacobi.
//
// Replaces SymmetricEigen::compute_jacobi / compute_gpu (src/eigen.rs:143-221, :309-355; GpuBackend::symmetric_eigen
// src/backends/gpu/mod.rs:466).  The reference's CPU path is CYCLIC Jacobi: pairs (p, q) in row order, one rotation
// after the other.  Rotations of disjoint pairs commute as far as the annihilated entries go, so the device runs the
// n/2 disjoint pairs of one round-robin round at once — n - 1 rounds cover every pair once, i.e. one "sweep" — with
// the reference's own formulas for one rotation (src/eigen.rs:248-306: tau, t, c, s; a_pp -= t a_pq, a_qq += t a_pq,
// a_pq = 0), the reference's skip rule (|a_pq| < 1e-7 * max(||A||_F, 1), src/eigen.rs:150-151, :166) and its
// stopping rule (a whole sweep without a rotation; at most 50 sweeps, else "Jacobi algorithm failed to converge after
// 50 sweeps").  The rotation ORDER differs from the CPU path, so results agree to rounding, not bit for bit; the
// parity tests state the tolerance (tests/test_eigen_gpu.py).
//
// One round = A' = J^T A J for the block-diagonal J of that round's rotations, and V' = V J:
//   jacobi_params_kernel   one thread per pair: (c, s, t) from the current A; per COLUMN j it leaves (partner, c_j,
//                          ss_j) with ss_j = -s for the first index of a pair, +s for the second, so that a rotated
//                          column is c_j * col_j + ss_j * col_partner for both roles
//   jacobi_apply_kernel    one CTA per pair: rows p and q of A staged in shared memory (coalesced), every new element
//                          of rows p and q from the four old ones it depends on,
//                              A'[i][j] = c2 * (c1 * a00 + ss1 * a10) + ss2 * (c1 * a01 + ss1 * a11),
//                          always evaluated with the SMALLER index as "1" — the transposed element is computed by
//                          another CTA from the same four numbers (A is symmetric) in the same order, so A stays
//                          EXACTLY symmetric without a mirroring pass; rows p, q of V^T rotated in the same kernel.
// A ping-pongs between two buffers; V is kept transposed (eigenvectors as rows: coalesced) and un-transposed, in
// eigenvalue order, by gather_columns_kernel at the end.  L2/HBM-bound: one read and one write of A per round (8 n^2
// bytes) plus the rotated rows of V^T; n - 1 rounds per sweep, typically 6-10 sweeps.
#include <algorithm>
#include <numeric>
#include <vector>

#include "common.cuh"

namespace trn {
namespace eig {

constexpr int kThreads = 256;

struct ColParam {      // how column (and, by symmetry, row) j is rotated this round
    float c, ss;       // new_j = c * old_j + ss * old_partner
    uint32_t partner;  // == j when j sits out (odd n) — then c = 1, ss = 0
    float tdelta;      // t * a_pq of j's pair: a_pp -= tdelta, a_qq += tdelta
    uint32_t rotated;  // 0: the pair was skipped (|a_pq| under the threshold) or j sits out
};

// round-robin ("circle") schedule over m = n rounded up to even players: pair 0 = (m - 1, r), pair i = (r + i, r - i)
__device__ __forceinline__ void round_pair(uint32_t m, uint32_t r, uint32_t i, uint32_t& p, uint32_t& q) {
    uint32_t a, b;
    if (i == 0) { a = m - 1; b = r; }
    else { a = (r + i) % (m - 1); b = (r + (m - 1) - i) % (m - 1); }
    p = min(a, b);
    q = max(a, b);
}

__global__ void jacobi_params_kernel(const float* __restrict__ a, uint32_t n, uint32_t m, uint32_t round, float tol,
                                     ColParam* __restrict__ prm, unsigned* __restrict__ rotations) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m / 2) return;
    uint32_t p, q;
    round_pair(m, round, i, p, q);
    if (q >= n) {   // p's partner is the padding player: p sits out
        if (p < n) prm[p] = ColParam{1.0f, 0.0f, p, 0.0f, 0u};
        return;
    }
    const float app = a[(size_t)p * n + p], aqq = a[(size_t)q * n + q], apq = a[(size_t)p * n + q];
    float c = 1.0f, s = 0.0f, td = 0.0f;
    if (!(fabsf(apq) < tol) && !(fabsf(apq) < 1e-15f)) {   // src/eigen.rs:166, :262
        const float tau = __fdiv_rn(__fsub_rn(aqq, app), __fmul_rn(2.0f, apq));
        const float root = __fsqrt_rn(__fadd_rn(1.0f, __fmul_rn(tau, tau)));
        const float t = tau >= 0.0f ? __fdiv_rn(1.0f, __fadd_rn(tau, root)) : __fdiv_rn(-1.0f, __fadd_rn(-tau, root));
        c = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(1.0f, __fmul_rn(t, t))));
        s = __fmul_rn(t, c);
        td = __fmul_rn(t, apq);
        atomicAdd(rotations, 1u);
        prm[p] = ColParam{c, -s, q, td, 1u};
        prm[q] = ColParam{c, s, p, td, 1u};
    } else {
        // skipped pair: identity rotation, the pair's 2x2 block stays as it is
        prm[p] = ColParam{1.0f, 0.0f, q, 0.0f, 0u};
        prm[q] = ColParam{1.0f, 0.0f, p, 0.0f, 0u};
    }
}

// c2 * (c1 * a00 + ss1 * a10) + ss2 * (c1 * a01 + ss1 * a11), every operation rounded on its own (no contraction):
// the transposed element evaluates the identical expression
__device__ __forceinline__ float rot2(float c1, float ss1, float c2, float ss2, float a00, float a10, float a01, float a11) {
    const float u = __fadd_rn(__fmul_rn(c1, a00), __fmul_rn(ss1, a10));
    const float w = __fadd_rn(__fmul_rn(c1, a01), __fmul_rn(ss1, a11));
    return __fadd_rn(__fmul_rn(c2, u), __fmul_rn(ss2, w));
}

__global__ void __launch_bounds__(kThreads)
jacobi_apply_kernel(const float* __restrict__ a_in, float* __restrict__ a_out, float* __restrict__ vt, uint32_t n, uint32_t m,
                    uint32_t round, const ColParam* __restrict__ prm) {
    extern __shared__ float rows[];   // [2][n]: row p, row q of the old A
    uint32_t p, q;
    round_pair(m, round, blockIdx.x, p, q);
    if (p >= n) return;
    const bool alone = q >= n;
    float* row_p = rows;
    float* row_q = rows + n;
    for (uint32_t j = threadIdx.x; j < n; j += kThreads) {
        row_p[j] = a_in[(size_t)p * n + j];
        row_q[j] = alone ? 0.0f : a_in[(size_t)q * n + j];
    }
    __syncthreads();
    const ColParam pp = prm[p];
    const ColParam pq = alone ? ColParam{1.0f, 0.0f, p, 0.0f, 0u} : prm[q];
    const bool rotated = pp.rotated != 0u;
    for (uint32_t j = threadIdx.x; j < n; j += kThreads) {
        const ColParam pj = prm[j];
        float new_p, new_q;
        if (!alone && (j == p || j == q)) {
            // the pair's own 2x2 block (src/eigen.rs:283-287); untouched when the pair was skipped
            if (rotated) {
                new_p = j == p ? __fsub_rn(row_p[p], pp.tdelta) : 0.0f;
                new_q = j == q ? __fadd_rn(row_q[q], pp.tdelta) : 0.0f;
            } else {
                new_p = row_p[j];
                new_q = row_q[j];
            }
        } else {
            const uint32_t jp = pj.partner;
            // row i of the pair against column j: "first" = the smaller of (i, j); old elements of rows i and partner(i)
            // row p: partner row is q
            if (p < j) new_p = rot2(pp.c, pp.ss, pj.c, pj.ss, row_p[j], row_q[j], row_p[jp], row_q[jp]);
            else       new_p = rot2(pj.c, pj.ss, pp.c, pp.ss, row_p[j], row_p[jp], row_q[j], row_q[jp]);
            if (!alone) {
                // row q: partner row is p
                if (q < j) new_q = rot2(pq.c, pq.ss, pj.c, pj.ss, row_q[j], row_p[j], row_q[jp], row_p[jp]);
                else       new_q = rot2(pj.c, pj.ss, pq.c, pq.ss, row_q[j], row_q[jp], row_p[j], row_p[jp]);
            } else {
                new_q = 0.0f;
            }
        }
        a_out[(size_t)p * n + j] = new_p;
        if (!alone) a_out[(size_t)q * n + j] = new_q;
    }
    // eigenvectors: columns p, q of V = rows p, q of V^T (src/eigen.rs:299-305)
    if (!alone && rotated) {
        const float c = pp.c, s = pq.ss;
        float* vp = vt + (size_t)p * n;
        float* vq = vt + (size_t)q * n;
        for (uint32_t k = threadIdx.x; k < n; k += kThreads) {
            const float x = vp[k], y = vq[k];
            vp[k] = __fsub_rn(__fmul_rn(c, x), __fmul_rn(s, y));
            vq[k] = __fadd_rn(__fmul_rn(s, x), __fmul_rn(c, y));
        }
    }
}

// working copy of A read through its UPPER triangle (the rotation decisions of src/eigen.rs:165 read a[i][j], i < j),
// so the copy is exactly symmetric whatever the caller's lower triangle holds; V^T = I
__global__ void init_kernel(const float* __restrict__ a, float* __restrict__ a0, float* __restrict__ vt, uint32_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * n) return;
    const uint32_t r = (uint32_t)(i / n), c = (uint32_t)(i % n);
    a0[i] = r <= c ? a[i] : a[(size_t)c * n + r];
    vt[i] = r == c ? 1.0f : 0.0f;
}

__global__ void diagonal_kernel(const float* __restrict__ a, uint32_t n, float* __restrict__ diag) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) diag[i] = a[(size_t)i * n + i];
}

// out[row][new_col] = vt[order[new_col]][row], eigenvalues[new_col] = diag[order[new_col]]: 32x32 tiles through shared memory
__global__ void gather_columns_kernel(const float* __restrict__ vt, const float* __restrict__ diag,
                                      const uint32_t* __restrict__ order, uint32_t n, float* __restrict__ vectors,
                                      float* __restrict__ values) {
    __shared__ float tile[32][33];
    const uint32_t r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    for (uint32_t y = threadIdx.y; y < 32; y += blockDim.y) {
        const uint32_t col = c0 + y, row = r0 + threadIdx.x;
        if (col < n && row < n) tile[y][threadIdx.x] = vt[(size_t)order[col] * n + row];
    }
    __syncthreads();
    for (uint32_t y = threadIdx.y; y < 32; y += blockDim.y) {
        const uint32_t row = r0 + y, col = c0 + threadIdx.x;
        if (row < n && col < n) vectors[(size_t)row * n + col] = tile[threadIdx.x][y];
    }
    if (blockIdx.x == 0 && threadIdx.y == 0) {
        const uint32_t col = c0 + threadIdx.x;
        if (col < n) values[col] = diag[order[col]];
    }
}

}  // namespace eig

size_t eigen_max_n() { return 8192; }   // two rows of A in shared memory (64 KiB) and 32-bit indexing

// a: n x n symmetric, row-major, on the device.  values[n] (descending) and vectors[n*n] (eigenvectors as columns) on
// the device.  Synchronises the stream once per sweep (the stopping rule is read back) and once for the sort.
int launch_symmetric_eigen(const float* a, size_t n_, float* values, float* vectors, int* sweeps_out, cudaStream_t s) {
    using namespace eig;
    Context* cx = ctx();
    if (!cx) return TRN_GPU_ERROR;
    const uint32_t n = (uint32_t)n_;
    const uint32_t m = (n + 1) & ~1u;
    const size_t nn = (size_t)n * n;
    float* scratch = nullptr;
    // A ping, A pong, V^T, diag[n], norm slot, params[n], order[n], rotation counter
    const size_t floats = 3 * nn + n + 4;
    const size_t bytes = floats * sizeof(float) + (size_t)n * sizeof(ColParam) + (size_t)n * sizeof(uint32_t) + 64;
    TRN_TRY(scratch_alloc((void**)&scratch, bytes, s));
    float* a0 = scratch;
    float* a1 = a0 + nn;
    float* vt = a1 + nn;
    float* diag = vt + nn;
    float* norm_slot = diag + n;
    ColParam* prm = reinterpret_cast<ColParam*>(norm_slot + 4);
    uint32_t* order = reinterpret_cast<uint32_t*>(prm + n);
    unsigned* counter = reinterpret_cast<unsigned*>(order + n);
    auto run = [&]() -> int {
        init_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, s>>>(a, a0, vt, n);
        count_launch();
        // tolerance = 1e-7 * max(||A||_F, 1)   (src/eigen.rs:150-151)
        TRN_TRY(launch_reduce(Reduce::NormL2, a0, nullptr, nn, norm_slot, s));
        float frob = 0.0f;
        TRN_CUDA(cudaMemcpyAsync(&frob, norm_slot, sizeof(float), cudaMemcpyDeviceToHost, s));
        TRN_CUDA(cudaStreamSynchronize(s));
        const float tol = 1e-7f * (frob > 1.0f ? frob : 1.0f);
        const size_t smem = 2 * (size_t)n * sizeof(float);
        if (smem > 48 * 1024) {
            static const cudaError_t optin = cudaFuncSetAttribute(jacobi_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            TRN_CUDA(optin);
        }
        float* cur = a0;
        float* nxt = a1;
        bool converged = n == 1;
        int sweep = 0;
        for (; sweep < 50 && !converged; ++sweep) {   // MAX_JACOBI_SWEEPS (src/eigen.rs:37)
            TRN_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), s));
            for (uint32_t r = 0; r + 1 < m; ++r) {
                jacobi_params_kernel<<<(m / 2 + 127) / 128, 128, 0, s>>>(cur, n, m, r, tol, prm, counter);
                jacobi_apply_kernel<<<m / 2, kThreads, smem, s>>>(cur, nxt, vt, n, m, r, prm);
                count_launch();
                count_launch();
                std::swap(cur, nxt);
            }
            TRN_CUDA(cudaGetLastError());
            unsigned rotations = 0;
            TRN_CUDA(cudaMemcpyAsync(&rotations, counter, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
            TRN_CUDA(cudaStreamSynchronize(s));
            converged = rotations == 0;
        }
        if (sweeps_out) *sweeps_out = sweep;
        if (!converged) return fail(TRN_INVALID_INPUT, "Jacobi algorithm failed to converge after 50 sweeps");
        // eigenvalues = diagonal, sorted descending with a stable sort (src/eigen.rs:183-190); columns permuted alike
        diagonal_kernel<<<(n + 255) / 256, 256, 0, s>>>(cur, n, diag);
        count_launch();
        std::vector<float> hdiag(n);
        TRN_CUDA(cudaMemcpyAsync(hdiag.data(), diag, n * sizeof(float), cudaMemcpyDeviceToHost, s));
        TRN_CUDA(cudaStreamSynchronize(s));
        std::vector<uint32_t> horder(n);
        std::iota(horder.begin(), horder.end(), 0u);
        std::stable_sort(horder.begin(), horder.end(), [&](uint32_t i, uint32_t j) { return hdiag[j] < hdiag[i]; });
        TRN_CUDA(cudaMemcpyAsync(order, horder.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        gather_columns_kernel<<<dim3((n + 31) / 32, (n + 31) / 32), dim3(32, 8), 0, s>>>(vt, diag, order, n, vectors, values);
        count_launch();
        TRN_CUDA(cudaGetLastError());
        TRN_CUDA(cudaStreamSynchronize(s));   // horder is read by the copy above
        return TRN_OK;
    };
    const int st = run();
    scratch_free(scratch, s);
    return st;
}

}  // namespace trn
