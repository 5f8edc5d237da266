// common.cuh — shared device helpers and the host-side context interface (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <mutex>

#include "trueno_cuda.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "trueno_b200 kernels are written for sm_100a only"
#endif

namespace trn {

// ---------------------------------------------------------------------------------------------
// Host-side context (context.cu)
// ---------------------------------------------------------------------------------------------
constexpr int kMaxReduceBlocks = 1 << 18;   // reduction grid upper bound = per-block partials in the workspace (3 MiB per stream)

struct Workspace {            // one per stream: scratch for the single-launch reductions
    float*    partial_val;    // [kMaxReduceBlocks] per-block partial values
    uint64_t* partial_idx;    // [kMaxReduceBlocks] per-block partial indices (arg reductions)
    unsigned* ticket;         // last-block-done counter; self-resetting
    float*    scalar_f32;     // device slot for host-slice scalar results
    uint64_t* scalar_u64;
    float*    host_f32;       // pinned mirrors of the two scalar slots
    uint64_t* host_u64;
};

struct Context {
    int device;
    int sm_count;
    int cc_major, cc_minor;
    size_t hbm_bytes;
    char name[256];
    cudaStream_t stream;      // the backend's own stream (host-slice calls, default for *_dev)
    cudaStream_t copy_stream; // H2D leg of the pipelined host-slice GEMM
    cudaStream_t d2h_stream;  // D2H leg of the pipelined host-slice GEMM
};

// Lazily binds the process to a device (LOCAL_RANK or 0); returns nullptr and sets the
// thread-local error if no sm_100 device is usable.  NEVER falls back to the CPU.
Context* ctx();
Workspace* workspace(cudaStream_t s);
cudaStream_t resolve_stream(void* s);
void count_launch(unsigned n = 1);
void uncount_launch(unsigned n);   // launchers called under stream capture record nodes, not launches

// thread-local error state -------------------------------------------------------------------
int fail(int status, const char* fmt, ...);
int fail_mismatch(size_t expected, size_t actual);
int fail_cuda(cudaError_t e, const char* what);

#define TRN_CUDA(expr)                                                     \
    do {                                                                   \
        cudaError_t _e = (expr);                                           \
        if (_e != cudaSuccess) return ::trn::fail_cuda(_e, #expr);         \
    } while (0)

#define TRN_TRY(expr)                                                      \
    do {                                                                   \
        int _s = (expr);                                                   \
        if (_s != TRN_OK) return _s;                                       \
    } while (0)

// stream-ordered scratch (cudaMallocAsync pool with an unlimited release threshold)
int scratch_alloc(void** p, size_t bytes, cudaStream_t s);
int scratch_free(void* p, cudaStream_t s);

// host <-> device transfers with pinned staging (context.cu)
int upload(float* dst_dev, const float* src_host, size_t n, cudaStream_t s);
int download(float* dst_host, const float* src_dev, size_t n, cudaStream_t s);

// ---------------------------------------------------------------------------------------------
// kernel launchers (one translation unit per family)
// ---------------------------------------------------------------------------------------------
// Peer-memory exchange context of a fused reduction (peer.cu builds it; world <= 1 disables the exchange).
constexpr int kMaxPeers = 8;
struct PeerCtx {
    unsigned long long* box[kMaxPeers];   // mailbox of every rank, mapped into this process (box[rank] is local)
    unsigned* seq;                        // DEVICE counter of the collective calls made on this communicator (lives beside
                                          // the local mailbox, bumped by the kernel itself: the calls can be captured in a
                                          // CUDA graph and replayed)
    unsigned* err;                        // host-mapped error words {call number, rank waited for + 1}: written by the first
                                          // exchange that gives up, read by the host entry points
    unsigned long long timeout_ns;        // spin budget of one exchange
    int rank, world;
};

enum class Reduce { Sum, Dot, SumSq, NormL2, SumAbs, MaxAbs, SumKahan };
int launch_reduce(Reduce op, const float* a, const float* b, size_t n, float* out, cudaStream_t s,
                  const PeerCtx* pc = nullptr);
// is_max: 1 argmax / 0 argmin.  out_idx / out_val may be null.  seed_rule 0: interior slice of a
// sharded vector (no a[0] seed; "no candidate" -> index ~0).
int launch_argreduce(int is_max, const float* a, size_t n, uint64_t* out_idx, float* out_val, cudaStream_t s,
                     int seed_rule = 1, uint64_t index_base = 0, trn_arg_pair* out_pair = nullptr,
                     const PeerCtx* pc = nullptr);
int launch_arg_combine(const trn_arg_pair* pairs, size_t count, int is_max, uint64_t* out_idx, float* out_val,
                       cudaStream_t s);

enum class Map { Add, Sub, Mul, Div, Scale, Abs, Clamp, Lerp, Fma, Relu, Exp, Sigmoid, Gelu, Swish, Tanh, Sqrt, Recip,
                 Ln, Log2, Log10, Sin, Cos, Tan, Floor, Ceil, Round,
                 // the rest of Vector's element-wise API (src/vector.rs:1448-4410)
                 Neg, Signum, Trunc, Fract, Sinh, Cosh, Asin, Acos, Atan, Asinh, Acosh, Atanh, Hardswish, Mish, Selu,
                 LeakyRelu, Elu, Pow, Minimum, Maximum, Copysign, Affine };
// out[i] = op(a[i], b[i], c[i]; p0, p1): b / c are read only by binary / ternary ops, p0 / p1 only by scale
// (p0 = scalar), clamp (p0 = min, p1 = max), lerp (p0 = t), leaky_relu (p0 = slope), elu (p0 = alpha), pow (p0 = n)
// and affine (out = (a - p0) * p1: zscore / minmax_normalize)
int launch_map(Map op, const float* a, const float* b, const float* c, float* out, size_t n, float p0, float p1,
               cudaStream_t s);

int launch_softmax_rows(int log_variant, const float* a, float* out, size_t rows, size_t cols, cudaStream_t s);
// one vector sharded over ranks: slice -> one (max, sum-of-exp relative to it) pair; then normalise the slice by the
// fold of every rank's pair (pairs: npairs x {max, sum}, rank order)
int launch_softmax_slice_stats(const float* a, size_t n, float* pair_out, cudaStream_t s);
int launch_softmax_slice_apply(const float* a, size_t n, const float* pairs, size_t npairs, int log_variant, float* out,
                               cudaStream_t s);

int launch_transpose(const float* a, size_t rows, size_t cols, float* out, cudaStream_t s);
// Matrix::embedding_lookup (gather.cu): out[r, :] = table[idx[r], :]; an index >= rows yields a zero row
int launch_gather_rows(const float* table, size_t rows, size_t cols, const uint64_t* idx, size_t n_idx, float* out,
                       cudaStream_t s);
int launch_convolve2d(const float* in, size_t rows, size_t cols, const float* kernel, size_t kr, size_t kc, float* out,
                      cudaStream_t s);
// fused attention (attention.cu): q, k, v, out [heads][seq][d]; engine 0 auto, 1 SIMT, 2 tcgen05
int launch_attention(const float* q, const float* k, const float* v, float* out, size_t heads, size_t seq, size_t d,
                     float scale, int causal, int engine, cudaStream_t s);
size_t attention_max_head_dim();
// SymmetricEigen (eigen.cu): parallel Jacobi; values descending, vectors as columns; synchronises the stream
int launch_symmetric_eigen(const float* a, size_t n, float* values, float* vectors, int* sweeps_out, cudaStream_t s);
size_t eigen_max_n();
int launch_matvec(const float* a, size_t rows, size_t cols, const float* v, float* y, cudaStream_t s);
int launch_vecmat(const float* x, const float* b, size_t k, size_t n, float* y, cudaStream_t s, bool skip_zero = true);
int launch_layer_norm_rows(const float* a, const float* gamma, const float* beta, float eps, float* out, size_t rows,
                           size_t cols, cudaStream_t s);
// C[b] = A[b] * B[b], b in [0,batch); strides in elements
// only_if_flag != nullptr: the kernel returns immediately unless *only_if_flag != 0 (device-side fallback)
int launch_gemm_simt(const float* a, const float* b, float* c, size_t batch, size_t m, size_t k, size_t n,
                     cudaStream_t s, const int* only_if_flag = nullptr);
// tcgen05 path; terms = 3 (3xTF32, fp32-accurate) or 1 (plain TF32, probe only)
// route_m != 0: `a` is a ROW BLOCK of a product with route_m rows; the kernel family is chosen as for the whole product
// (every kernel computes a row of C with the same arithmetic whatever tile it falls in, so the block's rows then carry
// the bits the unsharded product gives them)
int launch_gemm_tc(const float* a, const float* b, float* c, size_t batch, size_t m, size_t k, size_t n,
                   int terms, cudaStream_t s, size_t route_m = 0);
bool gemm_tc_supported(size_t m, size_t k, size_t n);
// the three phases of launch_gemm_tc, exposed so the host-slice path can pipeline them with transfers
size_t gemm_tc_kpad(size_t k);
int gemm_tc_split_a(const float* a, float* a_hi, float* a_lo, size_t batch, size_t m, size_t k, int* flag, cudaStream_t s);
int gemm_tc_split_b(const float* b, float* b_hi, float* b_lo, size_t batch, size_t k, size_t n, int* flag, cudaStream_t s);
// lo half only, for an operand whose raw words can serve as the hi half (the tensor core reads their top 19 bits)
bool gemm_tc_raw_hi_ok(const float* a, size_t k);
int gemm_tc_split_a_lo(const float* a, float* a_lo, size_t batch, size_t m, size_t k, int* flag, cudaStream_t s);
int gemm_tc_main(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo, float* c, size_t batch,
                 size_t m, size_t k, size_t n, int terms, const int* flag, cudaStream_t s, size_t route_m = 0);
// fused-split variant (no pre-pass, raw operands; gemm_tc.cu): preconditions + default rule, and the kernel itself
bool gemm_tc_uses_fused(const float* a, const float* b, size_t m, size_t k, size_t n);
int gemm_tc_fused_main(const float* a, const float* b, float* c, size_t batch, size_t m, size_t k, size_t n, int* flag,
                       cudaStream_t s);
// live timing brackets of one GEMM call (trn_profile_*): before the pre-pass, after it, after the main kernel
void gemm_profile_begin(cudaStream_t s);
void gemm_profile_mid(cudaStream_t s);
void gemm_profile_end(cudaStream_t s);
// auto-dispatch rule shared by the resident and the pipelined host paths (api.cu)
bool gemm_auto_uses_tc(size_t m, size_t k, size_t n);
bool is_pinned_host(const void* p);
// serialises the host-slice entry points, the command batch and trn_buf transfers on the backend's own stream
std::mutex& host_mutex();

// ---------------------------------------------------------------------------------------------
// Device helpers
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// 128-bit streaming load: read-only path, do not allocate in L1 (data is touched once)
__device__ __forceinline__ float4 ld_stream(const float4* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
// L2 eviction policies for two-pass kernels: pass 1 asks L2 to KEEP what it streams in (evict_last), pass 2 reads it
// back and lets it go (evict_first)
__device__ __forceinline__ uint64_t l2_policy_keep() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_drop() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ld_stream_hint(const float4* p, uint64_t policy) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(policy));
    return v;
}
// 128-bit streaming store (evict-first: the output is not re-read by this kernel)
__device__ __forceinline__ void st_stream(float4* p, const float4& v) {
    asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};"
                 :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream(float* p, float v) {
    asm volatile("st.global.cs.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ---- mbarrier / shared-memory PTX wrappers (TMA pipelines: gemm_tc.cu, softmax.cu) ----------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t"
        "}" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}

// 1-D bulk copy global -> shared through the TMA unit; completes `bytes` on the mbarrier.
// dst, src and bytes must be multiples of 16.
__device__ __forceinline__ void bulk_load_1d(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// named barrier over a subset of the CTA's warps (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t threads) {
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory");
}

#endif  // __CUDACC__

}  // namespace trn
