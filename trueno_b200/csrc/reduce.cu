// reduce.cu — single-launch, deterministic, HBM-streaming reductions for sm_100a.
//
// Replaces Avx2Backend::{dot,sum,max,min,argmax,argmin,norm_l2} (src/backends/avx2.rs:159-489)
// with the SEMANTICS of the scalar backend (src/backends/scalar.rs:66-200) — see DESIGN.md for why
// the AVX2 accumulation defects (single-accumulator sum, f32-lane indices) are not reproduced.
//
// Shape of every kernel:
//   * persistent grid of (SMs x resident CTAs) blocks, 256 threads; each block walks 16 KiB
//     (sum/arg) or 2x16 KiB (dot) tiles grid-strided, 4 independent 128-bit loads per thread per
//     array per step (ld.global.nc.L1::no_allocate) so >= 64 KiB/SM is in flight;
//   * per-thread accumulators -> warp shuffle tree -> shared-memory tree -> one partial per block;
//   * the LAST block to finish (ticket counter, self-resetting) folds the per-block partials in a
//     fixed order and writes the result.  No float atomics, fixed grid => bit-identical reruns
//     (the reference pins run-to-run determinism: tests/falsification_tests.rs:415-500).
//
// Algorithmic bytes per element: sum/max/min/argmax/argmin/norm_l2 4 B, dot 8 B; HBM-bound.
#include <cfloat>
#include <cstdlib>
#include <cmath>

#include "common.cuh"

namespace trn {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;                           // float4 loads in flight per thread per array
constexpr int kTileVec = kThreads * kUnroll;         // float4 per tile (16 KiB)

__device__ __forceinline__ bool last_block_done(unsigned* ticket) {
    __shared__ bool s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = atomicAdd(ticket, 1u);
        s_last = (t == gridDim.x - 1);
        if (s_last) *ticket = 0;  // self-reset for the next launch on this stream
    }
    __syncthreads();
    if (s_last) __threadfence();
    return s_last;
}

// ---- fused exchange over NVLink peer memory (slice partials + one exchange step) -------------
// Instead of a separate NCCL collective behind the kernel, the LAST block of every rank's reduction writes its
// slice result straight into every peer's mailbox (P2P stores through NVLink / NVSwitch) and then reads the
// results of all ranks from its own mailbox.  Payload and sequence number travel in ONE 64-bit store, so there
// is no flag/data ordering to get wrong; slots are double-buffered by call parity (a rank can only be one call
// ahead of its slowest peer, because finishing a call needs every peer's message of that call).  Every rank
// folds the values in rank order -> all ranks return the bit-identical result, run after run.
constexpr int kPeerWords = 4;   // 32-bit payload words per rank per call
__device__ __forceinline__ unsigned long long* peer_slot(unsigned long long* box, unsigned parity, int src, int word) {
    return box + ((parity * kMaxPeers + src) * kPeerWords + word);
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Returns false when a peer's message did not arrive within pc.timeout_ns: the payload of that rank is then NaN bits,
// the error words are raised for the host (the communicator is poisoned: later calls fail with GpuError before they
// launch) and the caller writes a NaN / "no index" result — a dead or desynchronised peer is an error, never a hang.
// `seq` = this call's number: the caller read (*pc.seq + 1) at KERNEL START (the counter only changes in the last block of a
// call, and calls on a communicator are stream-ordered), so the tail of the kernel does not wait for that load.
template <int NW>
__device__ __forceinline__ bool peer_exchange(const PeerCtx& pc, unsigned seq, const uint32_t (&mine)[NW], uint32_t (*all)[kPeerWords]) {
    __shared__ int s_failed;
    if (seq == 0) seq = 1;              // 0 is the "never written" state of a fresh mailbox
    if (threadIdx.x == 0) {
        *pc.seq = seq;
        s_failed = 0;
    }
    __syncthreads();
    const unsigned parity = seq & 1u;
    const int t = threadIdx.x;
    if (t < pc.world * NW) {
        const int dst = t / NW, w = t % NW;
        const unsigned long long msg = ((unsigned long long)seq << 32) | mine[w];
        // relaxed, not release: payload and call number are ONE word, nothing else is published through it (a
        // st.release.sys first drains every earlier write of the block to system scope: microseconds on the tail)
        asm volatile("st.relaxed.sys.global.u64 [%0], %1;" :: "l"(peer_slot(pc.box[dst], parity, pc.rank, w)), "l"(msg) : "memory");
    }
    if (t < pc.world * NW) {
        const int src = t / NW, w = t % NW;
        const unsigned long long* slot = peer_slot(pc.box[pc.rank], parity, src, w);
        const unsigned long long t0 = global_ns();
        unsigned long long v;
        unsigned spins = 0;
        bool ok = true;
        for (;;) {
            asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(slot) : "memory");
            if ((unsigned)(v >> 32) == seq) break;
            if ((++spins & 255u) == 0 && global_ns() - t0 > pc.timeout_ns) { ok = false; break; }
        }
        if (ok) {
            all[src][w] = (uint32_t)v;
        } else {
            all[src][w] = 0x7FC00000u;   // NaN
            if (atomicExch(&s_failed, 1) == 0) {
                volatile unsigned* e = pc.err;
                e[1] = (unsigned)src + 1u;
                __threadfence_system();
                e[0] = seq;
                __threadfence_system();
            }
        }
    }
    __syncthreads();
    return s_failed == 0;
}

__device__ __forceinline__ float block_sum(float v) {
    __shared__ float s_w[kThreads / 32];
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x < 32) {
        r = threadIdx.x < kThreads / 32 ? s_w[threadIdx.x] : 0.f;
        r = warp_sum(r);
    }
    __syncthreads();
    return r;  // valid in warp 0
}

__device__ __forceinline__ float block_max(float v) {
    __shared__ float s_m[kThreads / 32];
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    if (threadIdx.x < 32) {
        r = threadIdx.x < kThreads / 32 ? s_m[threadIdx.x] : 0.f;
        r = warp_max(r);
    }
    __syncthreads();
    return r;  // valid in warp 0
}

// ---- sum / dot / sum of squares / sum of |x| / max of |x| / compensated sum --------------------------------
// OP: 0 sum, 1 dot, 2 sumsq, 3 sum|x| (norm_l1, scalar.rs:218-228), 4 max|x| (norm_linf, scalar.rs:234-247:
//     starts at 0.0 and a NaN never wins, exactly fmaxf), 5 Kahan-compensated sum (sum_kahan, scalar.rs:170-183:
//     every thread runs the reference's compensation recurrence on its own elements; the per-thread
//     (sum, compensation) pairs then go through a fixed tree of error-free two-sums)
template <int OP>
__device__ __forceinline__ void accum1(float& acc, float& comp, float x, float y) {
    if (OP == 0) acc += x;
    else if (OP == 1) acc = fmaf(x, y, acc);
    else if (OP == 2) acc = fmaf(x, x, acc);
    else if (OP == 3) acc += fabsf(x);
    else if (OP == 4) acc = fmaxf(acc, fabsf(x));
    else {
        const float yk = __fsub_rn(x, comp);
        const float t = __fadd_rn(acc, yk);
        comp = __fsub_rn(__fsub_rn(t, acc), yk);
        acc = t;
    }
}
template <int OP>
__device__ __forceinline__ void accum4(float& acc, float& comp, const float4& x, const float4& y) {
    if (OP == 0) {
        acc += (x.x + x.y) + (x.z + x.w);
    } else if (OP == 1) {
        acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc);
        acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
    } else if (OP == 2) {
        acc = fmaf(x.x, x.x, acc); acc = fmaf(x.y, x.y, acc);
        acc = fmaf(x.z, x.z, acc); acc = fmaf(x.w, x.w, acc);
    } else if (OP == 3) {
        acc += (fabsf(x.x) + fabsf(x.y)) + (fabsf(x.z) + fabsf(x.w));
    } else if (OP == 4) {
        acc = fmaxf(acc, fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))));
    } else {
        accum1<OP>(acc, comp, x.x, 0.f); accum1<OP>(acc, comp, x.y, 0.f);
        accum1<OP>(acc, comp, x.z, 0.f); accum1<OP>(acc, comp, x.w, 0.f);
    }
}

// ---- compensated (hi, lo) pairs: the sum_kahan tree ---------------------------------------------------------
// (hi, lo) += (bh, bl) with Knuth's error-free two-sum: the rounding error of hi + bh moves into lo, so the
// cross-thread / cross-block tree keeps what the per-thread Kahan recurrences kept.
__device__ __forceinline__ void pair_add(float& hi, float& lo, float bh, float bl) {
    const float s = __fadd_rn(hi, bh);
    const float bb = __fsub_rn(s, hi);
    const float err = __fadd_rn(__fsub_rn(hi, __fsub_rn(s, bb)), __fsub_rn(bh, bb));
    hi = s;
    lo = __fadd_rn(__fadd_rn(lo, bl), err);
}
__device__ __forceinline__ void block_pair_sum(float& hi, float& lo) {   // result valid in thread 0
    __shared__ float s_h[kThreads / 32], s_l[kThreads / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float oh = __shfl_xor_sync(0xffffffffu, hi, o), ol = __shfl_xor_sync(0xffffffffu, lo, o);
        pair_add(hi, lo, oh, ol);
    }
    if ((threadIdx.x & 31) == 0) { s_h[threadIdx.x >> 5] = hi; s_l[threadIdx.x >> 5] = lo; }
    __syncthreads();
    if (threadIdx.x == 0) {
        hi = s_h[0]; lo = s_l[0];
#pragma unroll
        for (int w = 1; w < kThreads / 32; ++w) pair_add(hi, lo, s_h[w], s_l[w]);
    }
    __syncthreads();
}

// VEC: pointers are 16-byte aligned -> 128-bit path over n/4 vectors, scalar tail by block 0.
template <int OP, bool VEC, bool SQRT>
__global__ void __launch_bounds__(kThreads)
reduce_sum_kernel(const float* __restrict__ a, const float* __restrict__ b, size_t n,
                  float* __restrict__ partial, unsigned* __restrict__ ticket, float* __restrict__ out,
                  float* __restrict__ partial_lo, const PeerCtx pc) {
    __shared__ uint32_t s_all[kMaxPeers][kPeerWords];
    const unsigned call_seq = pc.world > 1 ? *pc.seq + 1u : 0u;
    float acc[kUnroll], comp[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) acc[u] = comp[u] = 0.f;

    if (VEC) {
        const size_t nvec = n >> 2;
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const float4* b4 = reinterpret_cast<const float4*>(b);
        // whole ROUNDS of 16 KiB tiles grid-strided (every block the same number), then the rest — up to one round of
        // tiles plus the ragged end — dealt in 4 KiB rows, so no block carries a whole tile more than its neighbour: at
        // 2^27 elements (one GPU's slice of 2^30 over 8) the uneven last round cost 3.6 % of the kernel
        const size_t full_tiles = (nvec / kTileVec) / gridDim.x * gridDim.x;
        for (size_t t = blockIdx.x; t < full_tiles; t += gridDim.x) {
            const size_t base = t * kTileVec + threadIdx.x;
            float4 x[kUnroll], y[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(a4 + base + u * kThreads);
            if (OP == 1) {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) y[u] = ld_stream(b4 + base + u * kThreads);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) accum4<OP>(acc[u], comp[u], x[u], OP == 1 ? y[u] : x[u]);
        }
        const size_t stride = (size_t)gridDim.x * kThreads;
        size_t v = full_tiles * kTileVec + (size_t)blockIdx.x * kThreads + threadIdx.x;
        for (; v + (kUnroll - 1) * stride < nvec; v += kUnroll * stride) {
            float4 x[kUnroll], y[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(a4 + v + u * stride);
            if (OP == 1) {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) y[u] = ld_stream(b4 + v + u * stride);
            }
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) accum4<OP>(acc[u], comp[u], x[u], OP == 1 ? y[u] : x[u]);
        }
        for (; v < nvec; v += stride) {
            float4 x = ld_stream(a4 + v);
            float4 y = OP == 1 ? ld_stream(b4 + v) : x;
            accum4<OP>(acc[0], comp[0], x, y);
        }
        if (blockIdx.x == 0) {
            size_t i = (nvec << 2) + threadIdx.x;
            if (i < n) accum1<OP>(acc[1], comp[1], a[i], OP == 1 ? b[i] : 0.f);
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads)
            accum1<OP>(acc[0], comp[0], ld_stream(a + i), OP == 1 ? ld_stream(b + i) : 0.f);
    }

    if (OP == 5) {
        // Kahan's c is the NEGATIVE of the part the running sum lost: thread total = acc - comp
        float hi = acc[0], lo = -comp[0];
#pragma unroll
        for (int u = 1; u < kUnroll; ++u) pair_add(hi, lo, acc[u], -comp[u]);
        block_pair_sum(hi, lo);
        if (threadIdx.x == 0) { partial[blockIdx.x] = hi; partial_lo[blockIdx.x] = lo; }
        if (last_block_done(ticket)) {
            float rh = 0.f, rl = 0.f;
            for (unsigned i = threadIdx.x; i < gridDim.x; i += kThreads) pair_add(rh, rl, __ldcg(partial + i), __ldcg(partial_lo + i));
            block_pair_sum(rh, rl);
            if (pc.world > 1) {   // (hi, lo) pairs of every slice, folded in rank order
                __shared__ float s_pair[2];
                if (threadIdx.x == 0) { s_pair[0] = rh; s_pair[1] = rl; }
                __syncthreads();
                const uint32_t mine[2] = {__float_as_uint(s_pair[0]), __float_as_uint(s_pair[1])};
                const bool ok = peer_exchange<2>(pc, call_seq, mine, s_all);
                rh = 0.f; rl = 0.f;
                for (int r = 0; r < pc.world; ++r) pair_add(rh, rl, __uint_as_float(s_all[r][0]), __uint_as_float(s_all[r][1]));
                if (!ok) rh = __uint_as_float(0x7FC00000u);
            }
            if (threadIdx.x == 0) *out = __fadd_rn(rh, rl);
        }
        return;
    }
    constexpr bool ISMAX = OP == 4;
    float v = ISMAX ? fmaxf(fmaxf(acc[0], acc[1]), fmaxf(acc[2], acc[3])) : (acc[0] + acc[1]) + (acc[2] + acc[3]);
    v = ISMAX ? block_max(v) : block_sum(v);
    if (threadIdx.x == 0) partial[blockIdx.x] = v;

    if (last_block_done(ticket)) {
        // fixed-order fold of the per-block partials (gridDim.x <= kMaxReduceBlocks)
        // (four independent chains per thread: a flat grid leaves up to 2^18 partials, and one dependent chain of L2 loads
        // per thread would put microseconds on the tail)
        float r4[4] = {0.f, 0.f, 0.f, 0.f};
        for (unsigned i = threadIdx.x; i < gridDim.x; i += 4 * kThreads) {
            float q[4];   // 0 is the identity of both folds (max|x| >= 0); all four loads are in flight together
#pragma unroll
            for (int u = 0; u < 4; ++u) q[u] = i + u * kThreads < gridDim.x ? __ldcg(partial + i + u * kThreads) : 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u) r4[u] = ISMAX ? fmaxf(r4[u], q[u]) : r4[u] + q[u];
        }
        float r = ISMAX ? fmaxf(fmaxf(r4[0], r4[1]), fmaxf(r4[2], r4[3])) : (r4[0] + r4[1]) + (r4[2] + r4[3]);
        r = ISMAX ? block_max(r) : block_sum(r);
        if (pc.world > 1) {   // slice totals of every rank, folded in rank order
            __shared__ float s_tot;
            if (threadIdx.x == 0) s_tot = r;
            __syncthreads();
            const uint32_t mine[1] = {__float_as_uint(s_tot)};
            const bool ok = peer_exchange<1>(pc, call_seq, mine, s_all);
            r = 0.f;
            for (int q = 0; q < pc.world; ++q) r = ISMAX ? fmaxf(r, __uint_as_float(s_all[q][0])) : r + __uint_as_float(s_all[q][0]);
            if (!ok) r = __uint_as_float(0x7FC00000u);
        }
        if (threadIdx.x == 0) *out = SQRT ? sqrtf(r) : r;
    }
}

// ---- argmax / argmin (also serves max / min: value at the winning index) ------------------------------
// Scalar-backend rule (src/backends/scalar.rs:112-166): seed with a[0], strict compare, first
// occurrence.  A NaN element never wins; a NaN seed never loses.  Within a thread indices are
// visited in increasing order, so a strict compare keeps the first occurrence; across threads the
// combine prefers the better value, then the lower index.
struct Best {
    float v;
    uint64_t i;
};
constexpr uint64_t kNoIndex = ~0ull;

template <bool MAX>
__device__ __forceinline__ bool better(float x, float v) { return MAX ? (x > v) : (x < v); }

template <bool MAX>
__device__ __forceinline__ Best combine(Best p, Best q) {
    if (better<MAX>(q.v, p.v) || (q.v == p.v && q.i < p.i)) return q;
    return p;
}
template <bool MAX>
__device__ __forceinline__ Best warp_best(Best p) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Best q;
        q.v = __shfl_xor_sync(0xffffffffu, p.v, o);
        q.i = __shfl_xor_sync(0xffffffffu, p.i, o);
        p = combine<MAX>(p, q);
    }
    return p;
}
template <bool MAX>
__device__ __forceinline__ Best block_best(Best p) {
    __shared__ float s_v[kThreads / 32];
    __shared__ uint64_t s_i[kThreads / 32];
    p = warp_best<MAX>(p);
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = p.v; s_i[threadIdx.x >> 5] = p.i; }
    __syncthreads();
    Best r{MAX ? -INFINITY : INFINITY, kNoIndex};
    if (threadIdx.x < 32) {
        if (threadIdx.x < kThreads / 32) { r.v = s_v[threadIdx.x]; r.i = s_i[threadIdx.x]; }
        r = warp_best<MAX>(r);
    }
    __syncthreads();
    return r;
}

// Cross-slice rule (slice 0 first): a NaN from slice 0 is the NaN seed and wins outright; otherwise best value,
// then LOWEST global index; NaN values and "no candidate" entries never win (src/backends/scalar.rs:140-166).
__device__ __forceinline__ bool pair_wins(int is_max, float v, uint64_t ix, float bv, uint64_t bi) {
    if (ix == kNoIndex || v != v) return false;
    return bi == kNoIndex || (is_max ? v > bv : v < bv) || (v == bv && ix < bi);
}

template <bool MAX, bool VEC>
__global__ void __launch_bounds__(kThreads)
argreduce_kernel(const float* __restrict__ a, size_t n, float* __restrict__ partial_v,
                 uint64_t* __restrict__ partial_i, unsigned* __restrict__ ticket,
                 uint64_t* __restrict__ out_idx, float* __restrict__ out_val, int seed_rule,
                 uint64_t index_base, trn_arg_pair* __restrict__ out_pair, const PeerCtx pc) {
    __shared__ uint32_t s_all[kMaxPeers][kPeerWords];
    const unsigned call_seq = pc.world > 1 ? *pc.seq + 1u : 0u;
    // a[0] seeds the scan (seed rule); read here, not by the last block behind the fold, where its latency would be serial
    const float seed0 = (seed_rule && n > 0 && threadIdx.x == 0) ? a[0] : 0.f;
    Best best{MAX ? -INFINITY : INFINITY, kNoIndex};
    auto visit = [&](float x, uint64_t i) { if (better<MAX>(x, best.v)) { best.v = x; best.i = i; } };

    if (VEC) {
        const size_t nvec = n >> 2;
        const float4* a4 = reinterpret_cast<const float4*>(a);
        const size_t full_tiles = (nvec / kTileVec) / gridDim.x * gridDim.x;   // whole rounds only, as in reduce_sum_kernel
        for (size_t t = blockIdx.x; t < full_tiles; t += gridDim.x) {
            const size_t base = t * kTileVec + threadIdx.x;
            float4 x[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(a4 + base + u * kThreads);
            // fast path: one min/max tree over the 16 values (fmaxf/fminf drop NaNs, like the strict compare)
            // and ONE test against the running best; the index bookkeeping runs only when a new best
            // appears, which becomes rare after the first few tiles.
            float m = MAX ? -INFINITY : INFINITY;
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {
                const float m2 = MAX ? fmaxf(fmaxf(x[u].x, x[u].y), fmaxf(x[u].z, x[u].w))
                                     : fminf(fminf(x[u].x, x[u].y), fminf(x[u].z, x[u].w));
                m = MAX ? fmaxf(m, m2) : fminf(m, m2);
            }
            if (better<MAX>(m, best.v)) {
#pragma unroll
                for (int u = 0; u < kUnroll; ++u) {   // ascending index order: the first occurrence wins
                    const uint64_t e = (uint64_t)(base + u * kThreads) << 2;
                    visit(x[u].x, e); visit(x[u].y, e + 1); visit(x[u].z, e + 2); visit(x[u].w, e + 3);
                }
            }
        }
        const size_t stride = (size_t)gridDim.x * kThreads;
        size_t v = full_tiles * kTileVec + (size_t)blockIdx.x * kThreads + threadIdx.x;
        for (; v + (kUnroll - 1) * stride < nvec; v += kUnroll * stride) {   // the last, partial round in 4 KiB rows
            float4 x[kUnroll];
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) x[u] = ld_stream(a4 + v + u * stride);
#pragma unroll
            for (int u = 0; u < kUnroll; ++u) {   // ascending index order within the thread: the first occurrence wins
                const uint64_t e = (uint64_t)(v + u * stride) << 2;
                visit(x[u].x, e); visit(x[u].y, e + 1); visit(x[u].z, e + 2); visit(x[u].w, e + 3);
            }
        }
        for (; v < nvec; v += stride) {
            float4 x = ld_stream(a4 + v);
            const uint64_t e = (uint64_t)v << 2;
            visit(x.x, e); visit(x.y, e + 1); visit(x.z, e + 2); visit(x.w, e + 3);
        }
        if (blockIdx.x == 0) {
            size_t i = (nvec << 2) + threadIdx.x;
            if (i < n) visit(a[i], i);
        }
    } else {
        for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < n; i += (size_t)gridDim.x * kThreads)
            visit(ld_stream(a + i), i);
    }

    best = block_best<MAX>(best);
    if (threadIdx.x == 0) { partial_v[blockIdx.x] = best.v; partial_i[blockIdx.x] = best.i; }

    if (last_block_done(ticket)) {
        Best r{MAX ? -INFINITY : INFINITY, kNoIndex};
        for (unsigned i = threadIdx.x; i < gridDim.x; i += 4 * kThreads) {   // four partials in flight per thread
            Best q[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                q[u] = Best{MAX ? -INFINITY : INFINITY, kNoIndex};
                if (i + u * kThreads < gridDim.x) q[u] = Best{__ldcg(partial_v + i + u * kThreads), __ldcg(partial_i + i + u * kThreads)};
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) r = combine<MAX>(r, q[u]);
        }
        r = block_best<MAX>(r);
        if (pc.world > 1) {
            // fused cross-slice pick: exchange (value, global index) with every rank, then apply the rule
            __shared__ float s_v;
            __shared__ uint64_t s_i;
            if (threadIdx.x == 0) {
                if (seed_rule && n > 0) {
                    const float seed = seed0;
                    if (seed != seed || r.i == kNoIndex) { r.v = seed; r.i = 0; }
                }
                s_v = r.v;
                s_i = r.i == kNoIndex ? kNoIndex : r.i + index_base;
            }
            __syncthreads();
            const uint32_t mine[3] = {__float_as_uint(s_v), (uint32_t)s_i, (uint32_t)(s_i >> 32)};
            const bool ok = peer_exchange<3>(pc, call_seq, mine, s_all);
            if (threadIdx.x == 0 && !ok) {
                if (out_idx) *out_idx = kNoIndex;
                if (out_val) *out_val = __uint_as_float(0x7FC00000u);
            }
            if (threadIdx.x == 0 && ok) {
                const float v0 = __uint_as_float(s_all[0][0]);
                float bv = MAX ? -INFINITY : INFINITY;
                uint64_t bi = kNoIndex;
                for (int q = 0; q < pc.world; ++q) {
                    const float v = __uint_as_float(s_all[q][0]);
                    const uint64_t ix = (uint64_t)s_all[q][1] | ((uint64_t)s_all[q][2] << 32);
                    if (pair_wins(MAX ? 1 : 0, v, ix, bv, bi)) { bv = v; bi = ix; }
                }
                if (v0 != v0) { bv = v0; bi = (uint64_t)s_all[0][1] | ((uint64_t)s_all[0][2] << 32); }
                if (out_idx) *out_idx = bi;
                if (out_val) *out_val = bv;
            }
            return;
        }
        if (threadIdx.x == 0) {
            // seed_rule == 1: this is a whole Vector (or its first slice) and a[0] seeds the scan.
            // A NaN seed never loses; and if nothing beat the identity, every element is the identity
            // value or NaN, so nothing is strictly better than a[0] either: the answer is index 0.
            // seed_rule == 0: an interior slice of a sharded vector — no seed; "no candidate" is
            // reported as index ~0 so the cross-slice combine can skip it (trueno_b200/parallel.py).
            if (seed_rule && n > 0) {
                const float seed = seed0;
                if (seed != seed || r.i == kNoIndex) { r.v = seed; r.i = 0; }
            }
            if (out_idx) *out_idx = r.i;
            if (out_val) *out_val = r.v;
            if (out_pair) {   // (value, GLOBAL index) for the cross-slice combine; "no candidate" stays ~0
                out_pair->value = r.v;
                out_pair->reserved = 0;
                out_pair->index = r.i == kNoIndex ? kNoIndex : r.i + index_base;
            }
        }
    }
}

// Cross-slice combine of `count` (value, global index) pairs (one per slice, slice 0 first) into the
// scalar-backend answer for the whole vector (src/backends/scalar.rs:140-166): a NaN from slice 0 is the
// NaN seed and wins outright; otherwise best value, then LOWEST global index; NaN values and "no
// candidate" entries never win.  One warp; replaces ~10 framework ops behind the all_gather.
__global__ void arg_combine_kernel(const trn_arg_pair* __restrict__ pairs, unsigned count, int is_max,
                                   uint64_t* __restrict__ out_idx, float* __restrict__ out_val) {
    const float seed_v = pairs[0].value;
    const bool nan_seed = seed_v != seed_v;
    float bv = is_max ? -INFINITY : INFINITY;
    uint64_t bi = kNoIndex;
    for (unsigned i = threadIdx.x; i < count; i += 32) {
        const float v = pairs[i].value;
        const uint64_t ix = pairs[i].index;
        if (ix == kNoIndex || v != v) continue;
        const bool win = bi == kNoIndex || (is_max ? v > bv : v < bv) || (v == bv && ix < bi);
        if (win) { bv = v; bi = ix; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const uint64_t oi = __shfl_xor_sync(0xffffffffu, bi, o);
        const bool win = oi != kNoIndex && (bi == kNoIndex || (is_max ? ov > bv : ov < bv) || (ov == bv && oi < bi));
        if (win) { bv = ov; bi = oi; }
    }
    if (threadIdx.x == 0) {
        if (nan_seed) { bv = seed_v; bi = pairs[0].index; }
        if (out_idx) *out_idx = bi;
        if (out_val) *out_val = bv;
    }
}

int launch_arg_combine(const trn_arg_pair* pairs, size_t count, int is_max, uint64_t* out_idx, float* out_val,
                       cudaStream_t s) {
    if (!ctx()) return TRN_GPU_ERROR;
    if (count == 0) return fail(TRN_INVALID_INPUT, "Empty vector");
    arg_combine_kernel<<<1, 32, 0, s>>>(pairs, (unsigned)count, is_max, out_idx, out_val);
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

// ---- launchers ------------------------------------------------------------------------------------------
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// enough blocks to cover the data once with 16 KiB tiles, capped at ONE resident wave (SMs x the
// CTAs of this kernel that really fit per SM — ncu showed 8/SM was 1.33 waves at 36 registers)
static int reduce_grid(size_t n, int sm_count, int per_sm) {
    size_t tiles = (n / 4 + kTileVec - 1) / kTileVec;
    size_t cap = (size_t)sm_count * (per_sm < 1 ? 1 : per_sm);
    if (cap > (size_t)kMaxReduceBlocks) cap = kMaxReduceBlocks;
    size_t g = tiles < cap ? tiles : cap;
    // TRN_REDUCE_FLAT = tiles per block of a FLAT grid (scripts/exp/exp_reduce_flat.py).  Measured slower at every size: a block's
    // fold + fence + ticket is paid per block, not once per SM slot (2^27 f32: 82.6 us persistent; 86.6 / 89.0 / 178.9 us at 16 /
    // 8 / 1 tiles per block).  Per-CTA timestamps of the persistent form (scripts/exp/exp_reduce_trace.cu, 2^27): the median CTA
    // ends its stream at 70.0 us (7.67 TB/s), the slowest at 73.7, the result is written at 75.5 and the next kernel's first CTA
    // starts 4.1 us later; dynamic chunk claiming and programmatic dependent launch changed none of it (79.8 -> 80.6 / 82.6 us).
    const char* e = getenv("TRN_REDUCE_FLAT");
    const int flat = e ? atoi(e) : 0;
    if (flat > 0) {
        size_t f = (tiles + flat - 1) / flat;
        if (f > (size_t)kMaxReduceBlocks) f = kMaxReduceBlocks;
        if (f > g) g = f;
    }
    return (int)(g ? g : 1);
}
template <class K>
static int blocks_per_sm(K kernel) {
    int v = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kernel, kThreads, 0) != cudaSuccess) { cudaGetLastError(); v = 4; }
    // Four resident CTAs per SM (64 KiB of loads in flight per array) stream as fast as the six to eight that fit, and the
    // smaller grid starts, drains and folds faster: sum over 2^30 f32 612 -> 586 us, over 2^27 (one GPU's slice at 8 GPUs)
    // 84.6 -> 80.1 us; two CTAs are too few (676 us).  TRN_REDUCE_PER_SM overrides (scripts/exp/exp_reduce_small.py).
    static const int cap = [] { const char* e = getenv("TRN_REDUCE_PER_SM"); const int c = e ? atoi(e) : 0; return c > 0 ? c : 4; }();
    v = v < 1 ? 1 : (v > 8 ? 8 : v);
    return cap < v ? cap : v;
}

int launch_reduce(Reduce op, const float* a, const float* b, size_t n, float* out, cudaStream_t s, const PeerCtx* pc) {
    Context* c = ctx();
    if (!c) return TRN_GPU_ERROR;
    PeerCtx pcv = {};
    if (pc) pcv = *pc;
    if (n == 0 && pcv.world <= 1) {  // empty -> 0.0 (src/vector.rs:635, :2602-2604; dot of empty slices is 0)
        TRN_CUDA(cudaMemsetAsync(out, 0, sizeof(float), s));
        return TRN_OK;
    }
    Workspace* w = workspace(s);
    if (!w) return fail(TRN_GPU_ERROR, "failed to allocate the reduction workspace");
    const bool vec = aligned16(a) && (op != Reduce::Dot || aligned16(b));
#define LAUNCH(OP, SQRT)                                                                                   \
    do {                                                                                                   \
        static const int per_sm = blocks_per_sm(reduce_sum_kernel<OP, true, SQRT>);   /* thread-safe, once */ \
        /* dot keeps two 16 KiB tiles in flight per CTA: three CTAs per SM stream best (2^30: 1205 -> 1191 us) */ \
        const int grid = reduce_grid(n, c->sm_count, (OP == 1 && per_sm > 3 && !getenv("TRN_REDUCE_PER_SM")) ? 3 : per_sm); \
        if (vec) reduce_sum_kernel<OP, true, SQRT><<<grid, kThreads, 0, s>>>(a, b, n, w->partial_val, w->ticket, out, reinterpret_cast<float*>(w->partial_idx), pcv); \
        else     reduce_sum_kernel<OP, false, SQRT><<<grid, kThreads, 0, s>>>(a, b, n, w->partial_val, w->ticket, out, reinterpret_cast<float*>(w->partial_idx), pcv); \
    } while (0)
    switch (op) {
        case Reduce::Sum:    LAUNCH(0, false); break;
        case Reduce::Dot:    LAUNCH(1, false); break;
        case Reduce::SumSq:  LAUNCH(2, false); break;
        case Reduce::NormL2: LAUNCH(2, true); break;
        case Reduce::SumAbs: LAUNCH(3, false); break;
        case Reduce::MaxAbs: LAUNCH(4, false); break;
        case Reduce::SumKahan: LAUNCH(5, false); break;
    }
#undef LAUNCH
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

int launch_argreduce(int is_max, const float* a, size_t n, uint64_t* out_idx, float* out_val, cudaStream_t s,
                     int seed_rule, uint64_t index_base, trn_arg_pair* out_pair, const PeerCtx* pc) {
    Context* c = ctx();
    if (!c) return TRN_GPU_ERROR;
    PeerCtx pcv = {};
    if (pc) pcv = *pc;
    Workspace* w = workspace(s);
    if (!w) return fail(TRN_GPU_ERROR, "failed to allocate the reduction workspace");
    static const int per_sm_max = blocks_per_sm(argreduce_kernel<true, true>);    // thread-safe, once
    static const int per_sm_min = blocks_per_sm(argreduce_kernel<false, true>);
    const int grid = reduce_grid(n, c->sm_count, is_max ? per_sm_max : per_sm_min);
    const bool vec = aligned16(a);
    if (is_max) {
        if (vec) argreduce_kernel<true, true><<<grid, kThreads, 0, s>>>(a, n, w->partial_val, w->partial_idx, w->ticket, out_idx, out_val, seed_rule, index_base, out_pair, pcv);
        else     argreduce_kernel<true, false><<<grid, kThreads, 0, s>>>(a, n, w->partial_val, w->partial_idx, w->ticket, out_idx, out_val, seed_rule, index_base, out_pair, pcv);
    } else {
        if (vec) argreduce_kernel<false, true><<<grid, kThreads, 0, s>>>(a, n, w->partial_val, w->partial_idx, w->ticket, out_idx, out_val, seed_rule, index_base, out_pair, pcv);
        else     argreduce_kernel<false, false><<<grid, kThreads, 0, s>>>(a, n, w->partial_val, w->partial_idx, w->ticket, out_idx, out_val, seed_rule, index_base, out_pair, pcv);
    }
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

}  // namespace trn
