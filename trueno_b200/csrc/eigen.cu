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
//   jacobi_round_kernel    ONE launch per round, one CTA per pair.  A round's rotations depend on A only through its
//                          diagonal and one off-diagonal element per pair; every CTA leaves those of its two rows
//                          for the NEXT round's pairing and the last CTA to finish (ticket) turns them into that
//                          round's per-COLUMN parameters (c_j, ss_j), ss_j = -s for the first index of a pair, +s
//                          for the second, so a rotated column is c_j * col_j + ss_j * col_partner in both roles.
//                          Rows p and q of A are staged in shared memory (coalesced), every new element
//                          of rows p and q comes from the four old ones it depends on,
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

// round-robin ("circle") schedule over m = n rounded up to even players, round r in [0, m - 1):
// pair 0 = (m - 1, r), pair i = (r + i, r - i) mod (m - 1); so the partner of index i is a closed form
__device__ __forceinline__ void round_pair(uint32_t m, uint32_t r, uint32_t i, uint32_t& p, uint32_t& q) {
    uint32_t a, b;
    if (i == 0) { a = m - 1; b = r; }
    else { a = (r + i) % (m - 1); b = (r + (m - 1) - i) % (m - 1); }
    p = min(a, b);
    q = max(a, b);
}
__device__ __forceinline__ uint32_t partner_of(uint32_t m, uint32_t r, uint32_t i) {
    if (i == m - 1) return r;
    if (i == r) return m - 1;
    return (2 * r + (m - 1) - i) % (m - 1);
}

// One rotation with the reference's expressions (src/eigen.rs:262-279); false when the pair is skipped
// (|a_pq| under the threshold, src/eigen.rs:166, or under 1e-15, :262)
__device__ __forceinline__ bool rotation(float app, float aqq, float apq, float tol, float& c, float& s, float& td) {
    c = 1.0f; s = 0.0f; td = 0.0f;
    if (fabsf(apq) < tol || fabsf(apq) < 1e-15f) return false;
    const float tau = __fdiv_rn(__fsub_rn(aqq, app), __fmul_rn(2.0f, apq));
    const float root = __fsqrt_rn(__fadd_rn(1.0f, __fmul_rn(tau, tau)));
    const float t = tau >= 0.0f ? __fdiv_rn(1.0f, __fadd_rn(tau, root)) : __fdiv_rn(-1.0f, __fadd_rn(-tau, root));
    c = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(1.0f, __fmul_rn(t, t))));
    s = __fmul_rn(t, c);
    td = __fmul_rn(t, apq);
    return true;
}

// c2 * (c1 * a00 + ss1 * a10) + ss2 * (c1 * a01 + ss1 * a11), every operation rounded on its own (no contraction):
// the transposed element evaluates the identical expression
__device__ __forceinline__ float rot2(float c1, float ss1, float c2, float ss2, float a00, float a10, float a01, float a11) {
    const float u = __fadd_rn(__fmul_rn(c1, a00), __fmul_rn(ss1, a10));
    const float w = __fadd_rn(__fmul_rn(c1, a01), __fmul_rn(ss1, a11));
    return __fadd_rn(__fmul_rn(c2, u), __fmul_rn(ss2, w));
}

struct ColParam {      // how column (and, by symmetry, row) j is rotated in a round: new_j = c * old_j + ss * old_partner
    float c, ss;       // ss = -s for the first index of a pair, +s for the second; (1, 0) when j's pair is skipped / j sits out
    float tdelta;      // t * a_pq of j's pair: a_pp -= tdelta, a_qq += tdelta
    uint32_t rotated;
};

// (c, ss) of every column for `round`, from the diagonal and offd[i] = A[i][partner(i, round)]; one thread per pair
__device__ __forceinline__ void pair_params(const volatile float* diag, const volatile float* offd, uint32_t n, uint32_t m,
                                            uint32_t round, uint32_t i, float tol, ColParam* __restrict__ prm) {
    uint32_t p, q;
    round_pair(m, round, i, p, q);
    if (p >= n) return;
    if (q >= n) { prm[p] = ColParam{1.0f, 0.0f, 0.0f, 0u}; return; }   // p sits out this round
    float c, s, td;
    const bool rot = rotation(diag[p], diag[q], offd[p], tol, c, s, td);
    prm[p] = ColParam{c, -s, td, rot ? 1u : 0u};
    prm[q] = ColParam{c, s, td, rot ? 1u : 0u};
}

__global__ void jacobi_first_params_kernel(const float* diag, const float* offd, uint32_t n, uint32_t m, float tol, ColParam* prm) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < m / 2) pair_params(diag, offd, n, m, 0, i, tol, prm);
}

// One round, ONE launch, one CTA per pair.  The rotations of a round depend on A only through its diagonal and one
// off-diagonal element per pair; every CTA leaves those of its two rows for the NEXT round's pairing (diag, offd[i] =
// A'[i][partner(i, next round)]), and the last CTA to finish (ticket) turns them into the next round's per-column
// parameters — so there is no separate parameter launch and no grid-wide wait inside the round.
__global__ void __launch_bounds__(kThreads)
jacobi_round_kernel(const float* __restrict__ a_in, float* __restrict__ a_out, float* __restrict__ vt,
                    const ColParam* __restrict__ prm_in, ColParam* __restrict__ prm_out, float* diag, float* offd,
                    unsigned* __restrict__ ticket, uint32_t n, uint32_t m, uint32_t round, float tol,
                    unsigned* __restrict__ rotations) {
    extern __shared__ float smem[];   // row p [n], row q [n] of the old A
    __shared__ bool s_last;
    float* row_p = smem;
    float* row_q = smem + n;
    uint32_t p, q;
    round_pair(m, round, blockIdx.x, p, q);
    const uint32_t next_round = round + 1 == m - 1 ? 0 : round + 1;
    if (p < n) {
        const bool alone = q >= n;
        for (uint32_t j = threadIdx.x; j < n; j += kThreads) {
            row_p[j] = a_in[(size_t)p * n + j];
            row_q[j] = alone ? 0.0f : a_in[(size_t)q * n + j];
        }
        __syncthreads();
        const ColParam pp = prm_in[p];
        const ColParam pq = alone ? ColParam{1.0f, 0.0f, 0.0f, 0u} : prm_in[q];
        const bool rotated = !alone && pp.rotated != 0u;
        if (rotated && threadIdx.x == 0) atomicAdd(rotations, 1u);
        const uint32_t next_p = partner_of(m, next_round, p), next_q = alone ? 0xFFFFFFFFu : partner_of(m, next_round, q);
        for (uint32_t j = threadIdx.x; j < n; j += kThreads) {
            float new_p, new_q = 0.0f;
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
                const ColParam pj = prm_in[j];
                const uint32_t jpart = partner_of(m, round, j);
                const uint32_t jq = jpart < n ? jpart : j;    // j sits out: its "partner" column is itself (c = 1, ss = 0)
                // row i of the pair against column j: "first" = the smaller of (i, j); old elements of rows i, partner(i)
                if (p < j) new_p = rot2(pp.c, pp.ss, pj.c, pj.ss, row_p[j], row_q[j], row_p[jq], row_q[jq]);
                else       new_p = rot2(pj.c, pj.ss, pp.c, pp.ss, row_p[j], row_p[jq], row_q[j], row_q[jq]);
                if (!alone) {
                    if (q < j) new_q = rot2(pq.c, pq.ss, pj.c, pj.ss, row_q[j], row_p[j], row_q[jq], row_p[jq]);
                    else       new_q = rot2(pj.c, pj.ss, pq.c, pq.ss, row_q[j], row_q[jq], row_p[j], row_p[jq]);
                }
            }
            a_out[(size_t)p * n + j] = new_p;
            if (j == p) diag[p] = new_p;
            if (j == next_p) offd[p] = new_p;
            if (!alone) {
                a_out[(size_t)q * n + j] = new_q;
                if (j == q) diag[q] = new_q;
                if (j == next_q) offd[q] = new_q;
            }
        }
        // eigenvectors: columns p, q of V = rows p, q of V^T (src/eigen.rs:299-305)
        if (rotated) {
            const float c = pp.c, sn = pq.ss;
            float* vp = vt + (size_t)p * n;
            float* vq = vt + (size_t)q * n;
            for (uint32_t k = threadIdx.x; k < n; k += kThreads) {
                const float x = vp[k], y = vq[k];
                vp[k] = __fsub_rn(__fmul_rn(c, x), __fmul_rn(sn, y));
                vq[k] = __fadd_rn(__fmul_rn(sn, x), __fmul_rn(c, y));
            }
        }
    }
    // last CTA of the round: the next round's parameters
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (s_last) {
        __threadfence();
        for (uint32_t i = threadIdx.x; i < m / 2; i += kThreads) pair_params(diag, offd, n, m, next_round, i, tol, prm_out);
        if (threadIdx.x == 0) *ticket = 0;
    }
}

// working copy of A read through its UPPER triangle (the rotation decisions of src/eigen.rs:165 read a[i][j], i < j),
// so the copy is exactly symmetric whatever the caller's lower triangle holds; V^T = I
__global__ void init_kernel(const float* __restrict__ a, float* __restrict__ a0, float* __restrict__ vt, float* __restrict__ diag,
                            float* __restrict__ offd, uint32_t n, uint32_t m) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)n * n) return;
    const uint32_t r = (uint32_t)(i / n), c = (uint32_t)(i % n);
    const float v = r <= c ? a[i] : a[(size_t)c * n + r];
    a0[i] = v;
    vt[i] = r == c ? 1.0f : 0.0f;
    if (r == c) diag[r] = v;
    if (c == partner_of(m, 0, r)) offd[r] = v;   // what round 0 needs
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
    // A ping, A pong, V^T, diag, offd, sorted-diag scratch, norm slot, params x2, order[n], rotation counter, ticket
    const size_t floats = 3 * nn + 3 * (size_t)n + 4;
    const size_t bytes = floats * sizeof(float) + 2 * (size_t)n * sizeof(ColParam) + (size_t)n * sizeof(uint32_t) + 64;
    TRN_TRY(scratch_alloc((void**)&scratch, bytes, s));
    float* a0 = scratch;
    float* a1 = a0 + nn;
    float* vt = a1 + nn;
    float* dg = vt + nn;
    float* od = dg + n;
    float* diag = od + n;
    float* norm_slot = diag + n;
    ColParam* prm[2] = {reinterpret_cast<ColParam*>(norm_slot + 4), reinterpret_cast<ColParam*>(norm_slot + 4) + n};
    uint32_t* order = reinterpret_cast<uint32_t*>(prm[1] + n);
    unsigned* counter = reinterpret_cast<unsigned*>(order + n);
    unsigned* ticket = counter + 1;
    auto run = [&]() -> int {
        TRN_CUDA(cudaMemsetAsync(counter, 0, 2 * sizeof(unsigned), s));
        init_kernel<<<(unsigned)((nn + 255) / 256), 256, 0, s>>>(a, a0, vt, dg, od, n, m);
        count_launch();
        // tolerance = 1e-7 * max(||A||_F, 1)   (src/eigen.rs:150-151)
        TRN_TRY(launch_reduce(Reduce::NormL2, a0, nullptr, nn, norm_slot, s));
        float frob = 0.0f;
        TRN_CUDA(cudaMemcpyAsync(&frob, norm_slot, sizeof(float), cudaMemcpyDeviceToHost, s));
        TRN_CUDA(cudaStreamSynchronize(s));
        const float tol = 1e-7f * (frob > 1.0f ? frob : 1.0f);
        jacobi_first_params_kernel<<<(m / 2 + 127) / 128, 128, 0, s>>>(dg, od, n, m, tol, prm[0]);
        count_launch();
        const size_t smem = 2 * (size_t)n * sizeof(float);
        if (smem > 48 * 1024) {
            static const cudaError_t optin = cudaFuncSetAttribute(jacobi_round_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            TRN_CUDA(optin);
        }
        float* cur = a0;
        float* nxt = a1;
        unsigned g = 0;   // rounds done: the parameter vectors ping-pong with it
        bool converged = n == 1;
        int sweep = 0;
        for (; sweep < 50 && !converged; ++sweep) {   // MAX_JACOBI_SWEEPS (src/eigen.rs:37)
            TRN_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned), s));
            for (uint32_t r = 0; r + 1 < m; ++r, ++g) {
                jacobi_round_kernel<<<m / 2, kThreads, smem, s>>>(cur, nxt, vt, prm[g & 1], prm[(g + 1) & 1], dg, od, ticket, n, m, r,
                                                                 tol, counter);
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
