// attention.cu — fused scaled-dot-product attention, O = softmax(scale * Q K^T [+ causal mask]) V, f32 in / f32 out.
//
// Replaces: trueno-gpu's `AttentionKernel` (trueno-gpu/src/kernels/attention.rs:27-125,
// parameters q/k/v/o [num_heads][seq_len][head_dim], `scale` = 1/sqrt(head_dim) by default, `causal`), i.e. the
// fusion of the composition the reference spells on the CPU as batched_matmul_4d(Q, K^T) (src/matrix.rs:464,
// the "attention pattern" test at :3985) -> Vector::scale -> Vector::softmax (src/vector.rs:1516) ->
// batched_matmul_4d(P, V).  The seq x seq score matrix never exists in HBM.
//
// attention_tf32x3_pair_kernel is the kernel that normally runs (two CTAs = two query tiles of a head, see below);
// attention_tf32x3_kernel is its single-CTA form (one query tile per head, short causal sequences, A/B runs).
// attention_tf32x3_kernel (head_dim <= 128): one CTA per (head, 128 query rows), 320 threads:
//   warp 0   TMA producer: the query tile (hi, lo) once, resident in shared memory; then a ring of 16 KiB stages
//            (5 at head_dim 128, up to 12 below) of 16-wide k-blocks of the pre-split K and V^T (hi, lo)
//   warp 1   MMA issuer (one elected lane), tcgen05.mma kind::tf32, 3xTF32 (lo*hi + hi*lo + hi*hi):
//              S  = Q K_j^T        A, B from shared memory            -> TMEM columns [0,128)
//              O' = P_j V_j        A = P (hi, lo) from TENSOR MEMORY  -> TMEM columns [384,512)
//   warps 2-9  softmax: thread = (query row, 64-column half).  Per 128-key tile: tcgen05.ld the scores, free the
//              S buffer at once (so S_{j+1} runs on the tensor pipe while this tile's exponentials run on the
//              CUDA cores), scale, mask, running max (halves exchange through shared memory), p = 2^(x - m),
//              split p into tf32 hi/lo and tcgen05.st them to TMEM columns [128,256) / [256,384); then drain the
//              previous tile's O' into register accumulators: O = O * exp(m_old - m_new) + O'.
//   Like the GEMM (gemm_tc.cu), TMEM only ever holds partial sums over 128 terms; the second accumulation level is
//   round-to-nearest in registers, which is also where the online-softmax rescaling happens — no TMEM rescale pass.
//   Issue order S_0, S_1, PV_0, S_2, PV_1, ...: the tensor pipe always has the next S queued behind a PV.
// attention_simt_kernel: one warp per query row, any head_dim <= 1024; the IEEE path for Inf/NaN inputs (gated on
//   the split pre-pass's device flag, as in the GEMM) and for head_dim > 128.
//
// Tensor-bound: 4 * seq^2 * head_dim flop per head (2 * seq^2 * head_dim for causal), executed 3x in TF32.
// HBM traffic is O(seq * head_dim) per head; K and V tiles are re-read from L2 by the seq/128 query tiles.
#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "tcgen05.cuh"

namespace trn {
namespace attn {

using namespace tc;

constexpr int BQ = 128;        // query rows per CTA (UMMA M)
constexpr int BKV = 128;       // keys per tile (UMMA N of S, K extent of one TMEM partial of O)
constexpr int SBK = 16;        // floats of K per ring stage (SWIZZLE_64B rows)
constexpr int kMaxStages = 12;
constexpr uint32_t kTile = 128 * SBK * 4;            // 8 KiB: 128 rows x 64 B
constexpr uint32_t kStageBytes = 2 * kTile;          // S phase: K_hi, K_lo; PV phase: V_hi, V_lo
constexpr uint32_t kQBlockBytes = 2 * kTile;         // one resident k-block of the query tile: Q_hi, Q_lo
constexpr int kThreads = 320;
constexpr int kPairThreads = 384;   // pair kernel: warpgroup 0 = TMA, MMA, 2 idle warps; warpgroups 1-2 = softmax
constexpr int kSoftmaxWarps = 8;
constexpr uint32_t kColS = 0, kColPhi = 128, kColPlo = 256, kColO = 384, kTmemCols = 512;
constexpr uint32_t kXchgBytes = 3 * 2 * 128 * 4;     // [buffer][half][row]
constexpr uint32_t kSmemBudget = 227 * 1024;
constexpr uint32_t kFixedBytes = kXchgBytes + 256 + 1024;   // exchange area, barriers, 1 KiB alignment slack
__host__ __device__ constexpr uint32_t ring_stages(uint32_t num_kb_s) {   // what is left after the resident Q tile
    const uint32_t n = (kSmemBudget - kFixedBytes - num_kb_s * kQBlockBytes) / kStageBytes;
    return n < (uint32_t)kMaxStages ? n : (uint32_t)kMaxStages;
}
constexpr uint32_t kHalfTile = kTile / 2;             // 64 rows x 64 B
constexpr uint32_t kPairStageBytes = 2 * kHalfTile;   // hi, lo of this CTA's half tile: 8 KiB
__host__ __device__ constexpr uint32_t pair_ring_stages(uint32_t num_kb_s) {
    const uint32_t n = (kSmemBudget - kFixedBytes - num_kb_s * kQBlockBytes) / kPairStageBytes;
    return n < (uint32_t)kMaxStages ? n : (uint32_t)kMaxStages;
}
__host__ __device__ constexpr uint32_t pair_smem_bytes(uint32_t num_kb_s) {
    return num_kb_s * kQBlockBytes + pair_ring_stages(num_kb_s) * kPairStageBytes + kFixedBytes;
}
__host__ __device__ constexpr uint32_t smem_bytes(uint32_t num_kb_s) {
    return num_kb_s * kQBlockBytes + ring_stages(num_kb_s) * kStageBytes + kFixedBytes;
}

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// round-to-nearest, ties away, to the 10-bit tf32 mantissa — cvt.rna.tf32.f32 for finite inputs (the PTX
// instruction is emulated with an extra Inf/NaN select on sm_100)
__device__ __forceinline__ uint32_t tf32_rna_finite(float v) { return (__float_as_uint(v) + 0x1000u) & 0xFFFFE000u; }

struct Params {
    float* out;
    const int* nonfinite_flag;
    uint32_t heads, seq, d;
    uint32_t dn;          // UMMA N of the PV product: head_dim rounded up to 16
    uint32_t num_kb_s;    // dpad / SBK
    uint32_t q_tiles;
    float scale;
    uint32_t causal;
};

__global__ void __launch_bounds__(kThreads, 1)
attention_tf32x3_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                        const __grid_constant__ CUtensorMap map_k_hi, const __grid_constant__ CUtensorMap map_k_lo,
                        const __grid_constant__ CUtensorMap map_v_hi, const __grid_constant__ CUtensorMap map_v_lo,
                        const Params p) {
    if (*p.nonfinite_flag != 0) return;   // Inf/NaN in the inputs: the SIMT kernel takes over (grid-uniform)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    // [Q tile: num_kb_s x (hi 8 KiB, lo 8 KiB), resident] [ring: kStages x (hi 8 KiB, lo 8 KiB)] [exchange] [barriers]
    const uint32_t kStages = ring_stages(p.num_kb_s);
    const uint32_t q_bytes = p.num_kb_s * kQBlockBytes;
    const uint32_t ring_base = smem_base + q_bytes;
    float* xchg = reinterpret_cast<float*>(smem_gen + q_bytes + kStages * kStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_gen + q_bytes + kStages * kStageBytes + kXchgBytes);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
    const uint32_t s_full = bar_base + 8u * (2 * kMaxStages), s_free = s_full + 8, p_full = s_full + 16,
                   o_full = s_full + 24, o_free = s_full + 32, q_full = s_full + 40;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 6);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // heavy (late, for causal) query tiles first
    const uint32_t head = blockIdx.x / p.q_tiles;
    const uint32_t qt = p.q_tiles - 1 - blockIdx.x % p.q_tiles;
    const uint32_t q0 = qt * BQ;
    const uint32_t kv_tiles = p.causal ? qt + 1 : (p.seq + BKV - 1) / BKV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q_hi); tma_prefetch_desc(&map_k_hi); tma_prefetch_desc(&map_v_hi);
        tma_prefetch_desc(&map_q_lo); tma_prefetch_desc(&map_k_lo); tma_prefetch_desc(&map_v_lo);
        for (uint32_t s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(q_full, 1);
        mbar_init(s_full, 1);
        mbar_init(s_free, kSoftmaxWarps);
        mbar_init(p_full, kSoftmaxWarps);
        mbar_init(o_full, 1);
        mbar_init(o_free, kSoftmaxWarps);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer: the MMA warp's order S_0, S_1, PV_0, S_2, PV_1, ... =====================
        if (elect_one()) {
            // the query tile, once: every k-block of Q_hi and Q_lo stays in shared memory for all key tiles
            mbar_expect_tx(q_full, q_bytes);
            for (uint32_t kb = 0; kb < p.num_kb_s; ++kb) {
                tma_load_3d(smem_base + kb * kQBlockBytes, &map_q_hi, q_full, (int)(kb * SBK), (int)q0, (int)head);
                tma_load_3d(smem_base + kb * kQBlockBytes + kTile, &map_q_lo, q_full, (int)(kb * SBK), (int)q0, (int)head);
            }
            uint32_t stage = 0, phase = 0;
            for (uint32_t step = 0; step <= kv_tiles; ++step) {
                if (step < kv_tiles) {
                    const int key0 = (int)(step * BKV);
                    for (uint32_t kb = 0; kb < p.num_kb_s; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        const uint32_t sa = ring_base + stage * kStageBytes;
                        mbar_expect_tx(full_bar(stage), 2 * kTile);
                        const int k0 = (int)(kb * SBK);
                        tma_load_3d(sa, &map_k_hi, full_bar(stage), k0, key0, (int)head);
                        tma_load_3d(sa + kTile, &map_k_lo, full_bar(stage), k0, key0, (int)head);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
                if (step >= 1) {
                    const int key0 = (int)((step - 1) * BKV);
                    for (uint32_t kb = 0; kb < BKV / SBK; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        const uint32_t sa = ring_base + stage * kStageBytes;
                        mbar_expect_tx(full_bar(stage), 2 * p.dn * SBK * 4);
                        tma_load_3d(sa, &map_v_hi, full_bar(stage), key0 + (int)(kb * SBK), 0, (int)head);
                        tma_load_3d(sa + kTile, &map_v_lo, full_bar(stage), key0 + (int)(kb * SBK), 0, (int)head);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc_s = make_idesc_tf32(BQ, BKV);
        const uint32_t idesc_o = make_idesc_tf32(BQ, (int)p.dn);
        // A k-block's six MMAs execute in 384 cycles, so the issue path between them must be a handful of instructions
        // (scripts/exp/exp_mma_rate.cu: 64 cycles per N = 128 MMA with descriptors at hand, 93 when each is rebuilt from
        // its address).  Every shared-memory address >> 4 fits the descriptor's 14-bit field, so a descriptor is a
        // constant high word and ONE add on the low word.
        constexpr uint32_t kDescHi = (uint32_t)(make_desc_k<SBK>(0) >> 32);
        constexpr uint32_t kT4 = kTile >> 4, kK4 = (UMMA_K * 4) >> 4;
        const uint32_t desc_q0 = (uint32_t)make_desc_k<SBK>(0) + (smem_base >> 4);      // resident Q tile
        const uint32_t desc_lo0 = (uint32_t)make_desc_k<SBK>(0) + (ring_base >> 4);     // ring
        auto desc = [](uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; };
        uint32_t stage = 0, phase = 0;
        mbar_wait(q_full, 0);
        for (uint32_t step = 0; step <= kv_tiles; ++step) {
            if (step < kv_tiles) {
                // ---- S_step = Q K^T over head_dim
                if (step >= 1) {
                    mbar_wait(s_free, (step - 1) & 1);   // the softmax warps have read S_{step-1}
                    tc_fence_after();
                }
                for (uint32_t kb = 0; kb < p.num_kb_s; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    if (elect_one()) {
                        // descriptor low words: one add each (see desc_lo0 above)
                        const uint32_t q_hi = desc_q0 + kb * (kQBlockBytes >> 4), q_lo = q_hi + kT4;
                        const uint32_t k_hi = desc_lo0 + stage * (kStageBytes >> 4), k_lo = k_hi + kT4;
                        const uint32_t dst = tmem_base + kColS;
#pragma unroll
                        for (int k = 0; k < SBK / UMMA_K; ++k) {
                            const uint32_t accum = (kb | (uint32_t)k) != 0;
                            umma_tf32(dst, desc(q_lo + k * kK4), desc(k_hi + k * kK4), idesc_s, accum);
                            umma_tf32(dst, desc(q_hi + k * kK4), desc(k_lo + k * kK4), idesc_s, 1u);
                            umma_tf32(dst, desc(q_hi + k * kK4), desc(k_hi + k * kK4), idesc_s, 1u);
                        }
                        umma_commit(empty_bar(stage));
                        if (kb == p.num_kb_s - 1) umma_commit(s_full);
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
            if (step >= 1) {
                // ---- O'_{t} = P_t V_t over the tile's 128 keys, t = step - 1
                const uint32_t t = step - 1;
                mbar_wait(p_full, t & 1);                      // P_t is in tensor memory
                if (t >= 1) mbar_wait(o_free, (t - 1) & 1);    // O'_{t-1} has been drained
                tc_fence_after();
                for (uint32_t kb = 0; kb < BKV / SBK; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t v_hi = desc_lo0 + stage * (kStageBytes >> 4), v_lo = v_hi + kT4;
                        const uint32_t dst = tmem_base + kColO;
                        const uint32_t p_hi = tmem_base + kColPhi + kb * SBK, p_lo = tmem_base + kColPlo + kb * SBK;
#pragma unroll
                        for (int k = 0; k < SBK / UMMA_K; ++k) {
                            const uint32_t accum = (kb | (uint32_t)k) != 0;
                            umma_tf32_ts(dst, p_lo + k * UMMA_K, desc(v_hi + k * kK4), idesc_o, accum);
                            umma_tf32_ts(dst, p_hi + k * UMMA_K, desc(v_lo + k * kK4), idesc_o, 1u);
                            umma_tf32_ts(dst, p_hi + k * UMMA_K, desc(v_hi + k * kK4), idesc_o, 1u);
                        }
                        umma_commit(empty_bar(stage));
                        if (kb == BKV / SBK - 1) umma_commit(o_full);   // O'_t complete; P_t no longer needed
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== softmax / accumulation warps =====================
        const uint32_t quad = warp & 3;                 // TMEM lane quadrant this warp may access
        const uint32_t half = (uint32_t)(warp - 2) >> 2;  // which 64 of the tile's 128 keys / of the output columns
        const uint32_t row_in_tile = quad * 32 + lane;
        const uint32_t qrow = q0 + row_in_tile;
        const uint32_t lane_addr = tmem_base + ((quad * 32u) << 16);
        const uint32_t ocols = p.dn > 64 * half ? min(64u, p.dn - 64 * half) : 0u;   // this thread's O columns
        float o_acc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) o_acc[i] = 0.0f;
        float m_run = -INFINITY, l_part = 0.0f, alpha_prev = 0.0f;
        const float scale2 = __fmul_rn(p.scale, 1.4426950408889634f);   // scale * log2(e)

        auto drain_o = [&](uint32_t t, float alpha) {
            // O = O * alpha + O'_t  (second accumulation level, round-to-nearest)
            mbar_wait(o_full, t & 1);
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if ((uint32_t)(j * 32) < ocols) {
                    uint32_t r[32];
                    tmem_ld_32x32(lane_addr + kColO + 64 * half + j * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o_acc[j * 32 + i] = __fmaf_rn(o_acc[j * 32 + i], alpha, __uint_as_float(r[i]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(o_free);
        };

        for (uint32_t t = 0; t < kv_tiles; ++t) {
            mbar_wait(s_full, t & 1);
            tc_fence_after();
            float x[64];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t r[32];
                tmem_ld_32x32(lane_addr + kColS + 64 * half + j * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) x[j * 32 + i] = __uint_as_float(r[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(s_free);   // S_{t+1} may overwrite the buffer

            // Scores in the base-2 domain: x = s * (scale * log2 e), so that exp(scale*s - m) = 2^(x - m2) is one FADD
            // and one MUFU.EX2 per element (ex2.approx: 2 ulp; the product rounding moves a score by <= 0.5 ulp —
            // both far inside the matmul contract's 1e-5 * sum|q||k| allowance on the scores).
            // Masking (keys past the sequence, keys after the query when causal) only on the tiles that need it.
            const uint32_t key0 = t * BKV + 64 * half;
            const bool mask_tile = (t + 1) * BKV > p.seq || (p.causal && t == qt);   // CTA-uniform
            float m_loc = -INFINITY;
            if (mask_tile) {
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    const uint32_t key = key0 + i;
                    const bool valid = key < p.seq && (!p.causal || key <= qrow);
                    x[i] = valid ? __fmul_rn(x[i], scale2) : -INFINITY;
                    m_loc = fmaxf(m_loc, x[i]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    x[i] = __fmul_rn(x[i], scale2);
                    m_loc = fmaxf(m_loc, x[i]);
                }
            }
            float* xb = xchg + (t & 1) * 256;
            xb[half * 128 + row_in_tile] = m_loc;
            named_bar_sync(1 + quad, 64);
            const float m_new = fmaxf(m_run, fmaxf(m_loc, xb[(half ^ 1) * 128 + row_in_tile]));
            const float m_use = m_new == -INFINITY ? 0.0f : m_new;            // row with no valid key yet
            const float alpha = m_run == -INFINITY ? 0.0f : ex2_approx(m_run - m_use);
            m_run = m_new;

            // p = 2^(x - m) in place, row-sum partial
            float l_tile = 0.0f;
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                x[i] = ex2_approx(x[i] - m_use);
                l_tile += x[i];
            }
            l_part = __fmaf_rn(l_part, alpha, l_tile);

            // the previous tile's PV product must have retired before P is overwritten; drain its O' first
            if (t >= 1) drain_o(t - 1, alpha_prev);
            alpha_prev = alpha;
            // tf32 split: hi = rna(p), lo = rna(p - hi)  (p - hi is exact), written where the MMA reads its A operand
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t r[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = tf32_rna_finite(x[j * 32 + i]);
                tmem_st_32x32(lane_addr + kColPhi + 64 * half + j * 32, r);
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = tf32_rna_finite(x[j * 32 + i] - __uint_as_float(r[i]));
                tmem_st_32x32(lane_addr + kColPlo + 64 * half + j * 32, r);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
        }
        drain_o(kv_tiles - 1, alpha_prev);

        // row sum: half 0 + half 1 (fixed order), then normalise and store this thread's columns
        float* xb = xchg + 2 * 256;
        xb[half * 128 + row_in_tile] = l_part;
        named_bar_sync(1 + quad, 64);
        const float inv_l = 1.0f / (xb[row_in_tile] + xb[128 + row_in_tile]);
        if (qrow < p.seq && ocols > 0) {
            float* orow = p.out + ((size_t)head * p.seq + qrow) * p.d + 64 * half;
            const uint32_t ncol = min(ocols, p.d > 64 * half ? p.d - 64 * half : 0u);
            const bool vec = (p.d & 3u) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15u) == 0;
#pragma unroll
            for (int i = 0; i < 64; i += 4) {
                if ((uint32_t)i + 3 < ncol && vec) {
                    *reinterpret_cast<float4*>(orow + i) = make_float4(o_acc[i] * inv_l, o_acc[i + 1] * inv_l, o_acc[i + 2] * inv_l, o_acc[i + 3] * inv_l);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if ((uint32_t)(i + e) < ncol) orow[i + e] = o_acc[i + e] * inv_l;
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---- CTA-pair variant (cta_group::2): the two CTAs of a cluster take two consecutive query tiles of one head and share
// every K and V^T tile — each loads HALF of it (64 keys, or 64 of V^T's rows) and the pair's MMAs (M = 256, issued by the
// leader) read both halves.  Per SM the score product then fetches 4 KiB of A + 2 KiB of B per MMA (96 B/clk instead of
// 128) and TMA writes half as much (21 B/clk instead of 43): under the shared-memory port's 128 B/clk, where the
// single-CTA kernel is over it; L2 -> SM traffic for K and V halves as well.  Synchronisation as in the CTA-pair GEMM:
// TMA loads of both CTAs complete_tx on the leader's barriers, tcgen05.commit multicasts "stage free" / "S ready" /
// "O' ready" to both CTAs, both CTAs' softmax warps arrive on the leader's s_free / p_full / o_free.
__global__ void __launch_bounds__(kPairThreads, 1)
attention_tf32x3_pair_kernel(const __grid_constant__ CUtensorMap map_q_hi, const __grid_constant__ CUtensorMap map_q_lo,
                        const __grid_constant__ CUtensorMap map_k_hi, const __grid_constant__ CUtensorMap map_k_lo,
                        const __grid_constant__ CUtensorMap map_v_hi, const __grid_constant__ CUtensorMap map_v_lo,
                        const Params p) {
    if (*p.nonfinite_flag != 0) return;   // Inf/NaN in the inputs: the SIMT kernel takes over (grid-uniform)
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    // [Q tile: num_kb_s x (hi 8 KiB, lo 8 KiB), resident] [ring: kStages x (hi 8 KiB, lo 8 KiB)] [exchange] [barriers]
    const uint32_t kStages = pair_ring_stages(p.num_kb_s);
    constexpr uint32_t kStageBytes = kPairStageBytes;   // this CTA's 64 of the tile's 128 keys (or 64 of V^T's rows): hi, lo
    const uint32_t q_bytes = p.num_kb_s * kQBlockBytes;
    const uint32_t ring_base = smem_base + q_bytes;
    float* xchg = reinterpret_cast<float*>(smem_gen + q_bytes + kStages * kStageBytes);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_gen + q_bytes + kStages * kStageBytes + kXchgBytes);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
    const uint32_t s_full = bar_base + 8u * (2 * kMaxStages), s_free = s_full + 8, p_full = s_full + 16,
                   o_full = s_full + 24, o_free = s_full + 32, q_full = s_full + 40;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 6);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    // one 2-CTA cluster per (head, PAIR of query tiles); heavy (late, for causal) pairs first; rank 0 leads
    const uint32_t rank = cluster_ctarank();
    const uint32_t q_pairs = (p.q_tiles + 1) / 2;
    const uint32_t pair_id = blockIdx.x >> 1;
    const uint32_t head = pair_id / q_pairs;
    const uint32_t qp = q_pairs - 1 - pair_id % q_pairs;
    const uint32_t qt = 2 * qp + rank;          // may be one past the last tile (odd tile count): rows masked, nothing stored
    const uint32_t q0 = qt * BQ;
    const uint32_t kv_tiles = p.causal ? min(2 * qp + 2, (p.seq + BKV - 1) / BKV) : (p.seq + BKV - 1) / BKV;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_q_hi); tma_prefetch_desc(&map_k_hi); tma_prefetch_desc(&map_v_hi);
        tma_prefetch_desc(&map_q_lo); tma_prefetch_desc(&map_k_lo); tma_prefetch_desc(&map_v_lo);
        for (uint32_t s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(q_full, 1);
        mbar_init(s_full, 1);
        mbar_init(s_free, 2 * kSoftmaxWarps);   // the leader's copies collect both CTAs' softmax warps
        mbar_init(p_full, 2 * kSoftmaxWarps);
        mbar_init(o_full, 1);
        mbar_init(o_free, 2 * kSoftmaxWarps);
        fence_barrier_init();
    }
    if (warp == 1) {   // the same warp in both CTAs allocates the pair's TMEM columns
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();     // barriers of BOTH CTAs are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Register budget by role (setmaxnreg is per warpgroup): warpgroup 0 = TMA warp, MMA warp and two idle warps gives
    // registers up; the softmax warpgroups take them — a softmax thread holds a 64-wide probability row AND a 64-wide
    // output accumulator, and at the uniform 168-register limit the accumulator spilled (the drain of O' took ~2700 cycles
    // per key tile and sat on the critical path between two value products).
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;" ::: "memory");
    if (warp == 0) {
        // ===================== TMA producer: the MMA warp's order S_0, S_1, PV_0, S_2, PV_1, ... =====================
        if (elect_one()) {
            // the query tile, once: every k-block of Q_hi and Q_lo stays in shared memory for all key tiles.
            // All loads of both CTAs complete_tx on the LEADER's barriers.
            const uint32_t lead_q = q_full & kPeerMask;
            if (rank == 0) mbar_expect_tx(q_full, 2 * q_bytes);
            for (uint32_t kb = 0; kb < p.num_kb_s; ++kb) {
                tma_load_3d_pair(smem_base + kb * kQBlockBytes, &map_q_hi, lead_q, (int)(kb * SBK), (int)q0, (int)head);
                tma_load_3d_pair(smem_base + kb * kQBlockBytes + kTile, &map_q_lo, lead_q, (int)(kb * SBK), (int)q0, (int)head);
            }
            uint32_t stage = 0, phase = 0;
            for (uint32_t step = 0; step <= kv_tiles; ++step) {
                if (step < kv_tiles) {
                    const int key0 = (int)(step * BKV + rank * (BKV / 2));   // this CTA's 64 keys of the tile
                    for (uint32_t kb = 0; kb < p.num_kb_s; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        const uint32_t sa = ring_base + stage * kStageBytes;
                        const uint32_t lead_full = full_bar(stage) & kPeerMask;
                        if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * kStageBytes);
                        const int k0 = (int)(kb * SBK);
                        tma_load_3d_pair(sa, &map_k_hi, lead_full, k0, key0, (int)head);
                        tma_load_3d_pair(sa + kHalfTile, &map_k_lo, lead_full, k0, key0, (int)head);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
                if (step >= 1) {
                    const int key0 = (int)((step - 1) * BKV);
                    const int row0 = (int)(rank * (p.dn / 2));               // this CTA's half of V^T's rows (output columns)
                    for (uint32_t kb = 0; kb < BKV / SBK; ++kb) {
                        mbar_wait(empty_bar(stage), phase ^ 1);
                        const uint32_t sa = ring_base + stage * kStageBytes;
                        const uint32_t lead_full = full_bar(stage) & kPeerMask;
                        if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * 2 * (p.dn / 2) * SBK * 4);
                        tma_load_3d_pair(sa, &map_v_hi, lead_full, key0 + (int)(kb * SBK), row0, (int)head);
                        tma_load_3d_pair(sa + kHalfTile, &map_v_lo, lead_full, key0 + (int)(kb * SBK), row0, (int)head);
                        if (++stage == kStages) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only): M = 256 = both CTAs' query tiles =====================
        if (rank == 0) {
        constexpr uint32_t idesc_s = make_idesc_tf32(2 * BQ, BKV);
        const uint32_t idesc_o = make_idesc_tf32(2 * BQ, (int)p.dn);
        // A k-block's six MMAs execute in 384 cycles, so the issue path between them must be a handful of instructions
        // (scripts/exp/exp_mma_rate.cu: 64 cycles per N = 128 MMA with descriptors at hand, 93 when each is rebuilt from
        // its address).  Every shared-memory address >> 4 fits the descriptor's 14-bit field, so a descriptor is a
        // constant high word and ONE add on the low word.
        constexpr uint32_t kDescHi = (uint32_t)(make_desc_k<SBK>(0) >> 32);
        constexpr uint32_t kT4 = kTile >> 4, kH4 = kHalfTile >> 4, kK4 = (UMMA_K * 4) >> 4;
        const uint32_t desc_q0 = (uint32_t)make_desc_k<SBK>(0) + (smem_base >> 4);      // resident Q tile
        const uint32_t desc_lo0 = (uint32_t)make_desc_k<SBK>(0) + (ring_base >> 4);     // ring
        auto desc = [](uint32_t lo) { return ((uint64_t)kDescHi << 32) | lo; };
        uint32_t stage = 0, phase = 0;
        mbar_wait(q_full, 0);
        for (uint32_t step = 0; step <= kv_tiles; ++step) {
            if (step < kv_tiles) {
                // ---- S_step = Q K^T over head_dim
                if (step >= 1) {
                    mbar_wait(s_free, (step - 1) & 1);   // the softmax warps have read S_{step-1}
                    tc_fence_after();
                }
                for (uint32_t kb = 0; kb < p.num_kb_s; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    if (elect_one()) {
                        // descriptor low words: one add each (see desc_lo0 above)
                        const uint32_t q_hi = desc_q0 + kb * (kQBlockBytes >> 4), q_lo = q_hi + kT4;
                        const uint32_t k_hi = desc_lo0 + stage * (kStageBytes >> 4), k_lo = k_hi + kH4;
                        const uint32_t dst = tmem_base + kColS;
#pragma unroll
                        for (int k = 0; k < SBK / UMMA_K; ++k) {
                            const uint32_t accum = (kb | (uint32_t)k) != 0;
                            umma_tf32_pair(dst, desc(q_lo + k * kK4), desc(k_hi + k * kK4), idesc_s, accum);
                            umma_tf32_pair(dst, desc(q_hi + k * kK4), desc(k_lo + k * kK4), idesc_s, 1u);
                            umma_tf32_pair(dst, desc(q_hi + k * kK4), desc(k_hi + k * kK4), idesc_s, 1u);
                        }
                        umma_commit_pair(empty_bar(stage));
                        if (kb == p.num_kb_s - 1) umma_commit_pair(s_full);
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
            if (step >= 1) {
                // ---- O'_{t} = P_t V_t over the tile's 128 keys, t = step - 1
                const uint32_t t = step - 1;
                mbar_wait(p_full, t & 1);                      // P_t is in tensor memory
                if (t >= 1) mbar_wait(o_free, (t - 1) & 1);    // O'_{t-1} has been drained
                tc_fence_after();
                for (uint32_t kb = 0; kb < BKV / SBK; ++kb) {
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t v_hi = desc_lo0 + stage * (kStageBytes >> 4), v_lo = v_hi + kH4;
                        const uint32_t dst = tmem_base + kColO;
                        const uint32_t p_hi = tmem_base + kColPhi + kb * SBK, p_lo = tmem_base + kColPlo + kb * SBK;
#pragma unroll
                        for (int k = 0; k < SBK / UMMA_K; ++k) {
                            const uint32_t accum = (kb | (uint32_t)k) != 0;
                            umma_tf32_ts_pair(dst, p_lo + k * UMMA_K, desc(v_hi + k * kK4), idesc_o, accum);
                            umma_tf32_ts_pair(dst, p_hi + k * UMMA_K, desc(v_lo + k * kK4), idesc_o, 1u);
                            umma_tf32_ts_pair(dst, p_hi + k * UMMA_K, desc(v_hi + k * kK4), idesc_o, 1u);
                        }
                        umma_commit_pair(empty_bar(stage));
                        if (kb == BKV / SBK - 1) umma_commit_pair(o_full);   // O'_t complete; P_t no longer needed
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
        }
    }
    } else {
        // ===================== softmax / accumulation warps (4..11) =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;" ::: "memory");
        const uint32_t quad = warp & 3;                 // TMEM lane quadrant this warp may access
        const uint32_t half = (uint32_t)(warp - 4) >> 2;  // which 64 of the tile's 128 keys / of the output columns
        const uint32_t row_in_tile = quad * 32 + lane;
        const uint32_t qrow = q0 + row_in_tile;
        const uint32_t lane_addr = tmem_base + ((quad * 32u) << 16);
        const uint32_t ocols = p.dn > 64 * half ? min(64u, p.dn - 64 * half) : 0u;   // this thread's O columns
        float o_acc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) o_acc[i] = 0.0f;
        float m_run = -INFINITY, l_part = 0.0f, alpha_prev = 0.0f;
        const float scale2 = __fmul_rn(p.scale, 1.4426950408889634f);   // scale * log2(e)

        auto drain_o = [&](uint32_t t, float alpha) {
            // O = O * alpha + O'_t  (second accumulation level, round-to-nearest)
            mbar_wait(o_full, t & 1);
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                if ((uint32_t)(j * 32) < ocols) {
                    uint32_t r[32];
                    tmem_ld_32x32(lane_addr + kColO + 64 * half + j * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o_acc[j * 32 + i] = __fmaf_rn(o_acc[j * 32 + i], alpha, __uint_as_float(r[i]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(o_free & kPeerMask);
        };

        for (uint32_t t = 0; t < kv_tiles; ++t) {
            mbar_wait(s_full, t & 1);
            tc_fence_after();
            float x[64];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t r[32];
                tmem_ld_32x32(lane_addr + kColS + 64 * half + j * 32, r);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) x[j * 32 + i] = __uint_as_float(r[i]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(s_free & kPeerMask);   // S_{t+1} may overwrite the buffer

            // Scores in the base-2 domain: x = s * (scale * log2 e), so that exp(scale*s - m) = 2^(x - m2) is one FADD
            // and one MUFU.EX2 per element (ex2.approx: 2 ulp; the product rounding moves a score by <= 0.5 ulp —
            // both far inside the matmul contract's 1e-5 * sum|q||k| allowance on the scores).
            // Masking (keys past the sequence, keys after the query when causal) only on the tiles that need it.
            const uint32_t key0 = t * BKV + 64 * half;
            const bool mask_tile = (t + 1) * BKV > p.seq || (p.causal && t >= qt);   // CTA-uniform; t > qt: wholly masked
            float m_loc = -INFINITY;
            if (mask_tile) {
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    const uint32_t key = key0 + i;
                    const bool valid = key < p.seq && (!p.causal || key <= qrow);
                    x[i] = valid ? __fmul_rn(x[i], scale2) : -INFINITY;
                    m_loc = fmaxf(m_loc, x[i]);
                }
            } else {
#pragma unroll
                for (int i = 0; i < 64; ++i) {
                    x[i] = __fmul_rn(x[i], scale2);
                    m_loc = fmaxf(m_loc, x[i]);
                }
            }
            float* xb = xchg + (t & 1) * 256;
            xb[half * 128 + row_in_tile] = m_loc;
            named_bar_sync(1 + quad, 64);
            const float m_new = fmaxf(m_run, fmaxf(m_loc, xb[(half ^ 1) * 128 + row_in_tile]));
            const float m_use = m_new == -INFINITY ? 0.0f : m_new;            // row with no valid key yet
            const float alpha = m_run == -INFINITY ? 0.0f : ex2_approx(m_run - m_use);
            m_run = m_new;

            // p = 2^(x - m) in place, row-sum partial
            float l_tile = 0.0f;
#pragma unroll
            for (int i = 0; i < 64; ++i) {
                x[i] = ex2_approx(x[i] - m_use);
                l_tile += x[i];
            }
            l_part = __fmaf_rn(l_part, alpha, l_tile);

            // the previous tile's PV product must have retired before P is overwritten; drain its O' first
            if (t >= 1) drain_o(t - 1, alpha_prev);
            alpha_prev = alpha;
            // tf32 split: hi = rna(p), lo = rna(p - hi)  (p - hi is exact), written where the MMA reads its A operand
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                uint32_t r[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = tf32_rna_finite(x[j * 32 + i]);
                tmem_st_32x32(lane_addr + kColPhi + 64 * half + j * 32, r);
#pragma unroll
                for (int i = 0; i < 32; ++i) r[i] = tf32_rna_finite(x[j * 32 + i] - __uint_as_float(r[i]));
                tmem_st_32x32(lane_addr + kColPlo + 64 * half + j * 32, r);
            }
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(p_full & kPeerMask);
        }
        drain_o(kv_tiles - 1, alpha_prev);

        // row sum: half 0 + half 1 (fixed order), then normalise and store this thread's columns
        float* xb = xchg + 2 * 256;
        xb[half * 128 + row_in_tile] = l_part;
        named_bar_sync(1 + quad, 64);
        const float inv_l = 1.0f / (xb[row_in_tile] + xb[128 + row_in_tile]);
        if (qrow < p.seq && ocols > 0) {
            float* orow = p.out + ((size_t)head * p.seq + qrow) * p.d + 64 * half;
            const uint32_t ncol = min(ocols, p.d > 64 * half ? p.d - 64 * half : 0u);
            const bool vec = (p.d & 3u) == 0 && (reinterpret_cast<uintptr_t>(p.out) & 15u) == 0;
#pragma unroll
            for (int i = 0; i < 64; i += 4) {
                if ((uint32_t)i + 3 < ncol && vec) {
                    *reinterpret_cast<float4*>(orow + i) = make_float4(o_acc[i] * inv_l, o_acc[i + 1] * inv_l, o_acc[i + 2] * inv_l, o_acc[i + 3] * inv_l);
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        if ((uint32_t)(i + e) < ncol) orow[i + e] = o_acc[i + e] * inv_l;
                }
            }
        }
    }

    // neither CTA may leave (or free TMEM) while its partner can still signal it or read its shared memory
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- IEEE path: one warp per query row, online softmax key by key -------------------------------------------
constexpr int kSimtMaxC = 32;   // head_dim <= 1024

__global__ void __launch_bounds__(128)
attention_simt_kernel(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                      float* __restrict__ out, uint32_t heads, uint32_t seq, uint32_t d, float scale, uint32_t causal,
                      const int* __restrict__ only_if_flag) {
    if (only_if_flag != nullptr && *only_if_flag == 0) return;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (size_t row_id = (size_t)blockIdx.x * 4 + warp; row_id < (size_t)heads * seq; row_id += (size_t)gridDim.x * 4) {
    const uint32_t head = (uint32_t)(row_id / seq), qrow = (uint32_t)(row_id % seq);
    const float* qr = q + row_id * d;
    const float* kh = k + (size_t)head * seq * d;
    const float* vh = v + (size_t)head * seq * d;
    float qv[kSimtMaxC], o[kSimtMaxC];
#pragma unroll
    for (int c = 0; c < kSimtMaxC; ++c) {
        const uint32_t col = c * 32 + lane;
        qv[c] = col < d ? qr[col] : 0.0f;
        o[c] = 0.0f;
    }
    const uint32_t nc = (d + 31) / 32;
    float m = -INFINITY, l = 0.0f;
    const uint32_t last = causal ? qrow + 1 : seq;
    for (uint32_t j = 0; j < last; ++j) {
        float part = 0.0f;
#pragma unroll
        for (int c = 0; c < kSimtMaxC; ++c) {
            if ((uint32_t)c < nc) {
                const uint32_t col = c * 32 + lane;
                if (col < d) part = fmaf(qv[c], kh[(size_t)j * d + col], part);
            }
        }
        const float x = warp_sum(part) * scale;
        const float m_new = fmaxf(m, x);
        // NaN scores propagate through fmaxf-free arithmetic below: keep m finite-or-NaN consistent
        const float alpha = m == -INFINITY ? 0.0f : expf(m - m_new);
        const float pv = expf(x - m_new);
        l = fmaf(l, alpha, pv);
#pragma unroll
        for (int c = 0; c < kSimtMaxC; ++c) {
            if ((uint32_t)c < nc) {
                const uint32_t col = c * 32 + lane;
                if (col < d) o[c] = fmaf(o[c], alpha, pv * vh[(size_t)j * d + col]);
            }
        }
        m = m_new;
    }
#pragma unroll
    for (int c = 0; c < kSimtMaxC; ++c) {
        const uint32_t col = c * 32 + lane;
        if ((uint32_t)c < nc && col < d) out[row_id * d + col] = o[c] / l;
    }
    }
}

}  // namespace attn

size_t attention_max_head_dim() { return (size_t)attn::kSimtMaxC * 32; }

// q, k, v, out: [heads][seq][d] row-major f32 on the device.  engine: 0 auto, 1 SIMT, 2 tensor cores.
int launch_attention(const float* q, const float* k, const float* v, float* out, size_t heads, size_t seq, size_t d,
                     float scale, int causal, int engine, cudaStream_t s) {
    using namespace attn;
    Context* cx = ctx();
    if (!cx) return TRN_GPU_ERROR;
    if (heads == 0 || seq == 0 || d == 0) return TRN_OK;
    const size_t rows = heads * seq;
    const size_t simt_blocks = (rows + 3) / 4;
    if (simt_blocks > 0x7FFFFFFFull || seq >= (1u << 30)) return fail(TRN_INVALID_INPUT, "attention over %zu rows exceeds the launch grid", rows);
    const bool tc_ok = d <= 128 && heads * ((seq + BQ - 1) / BQ) <= 0x7FFFFFFFull;
    if (engine == 2 && !tc_ok) return fail(TRN_INVALID_INPUT, "tensor-core attention needs head_dim <= 128, got %zu", d);
    if (engine == 1 || !tc_ok) {
        attention_simt_kernel<<<(unsigned)simt_blocks, 128, 0, s>>>(q, k, v, out, (uint32_t)heads, (uint32_t)seq, (uint32_t)d,
                                                                   scale, causal ? 1u : 0u, nullptr);
        count_launch();
        TRN_CUDA(cudaGetLastError());
        return TRN_OK;
    }

    // split pre-pass: Q, K -> hi/lo [heads][seq][dpad] (K-major as they are); V -> hi/lo [heads][d][seqpad] (transposed)
    const size_t dpad = gemm_tc_kpad(d), seqpad = gemm_tc_kpad(seq);
    const size_t qk_elems = heads * seq * dpad, v_elems = heads * d * seqpad;
    float* scratch = nullptr;
    TRN_TRY(scratch_alloc((void**)&scratch, (4 * qk_elems + 2 * v_elems) * sizeof(float) + 256, s));
    float* q_hi = scratch;
    float* q_lo = q_hi + qk_elems;
    float* k_hi = q_lo + qk_elems;
    float* k_lo = k_hi + qk_elems;
    float* v_hi = k_lo + qk_elems;
    float* v_lo = v_hi + v_elems;
    int* flag = reinterpret_cast<int*>(v_lo + v_elems);
    int st = TRN_OK;
    auto run = [&]() -> int {
        TRN_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));
        // Q and K are K-major as they lie ([seq][d]): when head_dim needs no padding their RAW words are the hi operands
        // (the tensor core reads the top 19 bits) and the pre-pass writes the lo halves only — 7 instead of 9 array passes
        // over q / k / v (TRN_ATT_RAW_HI=0 writes rounded hi halves as before).  V is transposed on the way, so both halves go out.
        static const bool raw_env = [] { const char* e = getenv("TRN_ATT_RAW_HI"); return !(e && atoi(e) == 0); }();
        const bool raw_q = raw_env && gemm_tc_raw_hi_ok(q, d), raw_k = raw_env && gemm_tc_raw_hi_ok(k, d);
        if (raw_q) { TRN_TRY(gemm_tc_split_a_lo(q, q_lo, heads, seq, d, flag, s)); q_hi = const_cast<float*>(q); }
        else TRN_TRY(gemm_tc_split_a(q, q_hi, q_lo, heads, seq, d, flag, s));
        if (raw_k) { TRN_TRY(gemm_tc_split_a_lo(k, k_lo, heads, seq, d, flag, s)); k_hi = const_cast<float*>(k); }
        else TRN_TRY(gemm_tc_split_a(k, k_hi, k_lo, heads, seq, d, flag, s));
        TRN_TRY(gemm_tc_split_b(v, v_hi, v_lo, heads, seq, d, flag, s));
        Params p;
        p.out = out;
        p.nonfinite_flag = flag;
        p.heads = (uint32_t)heads;
        p.seq = (uint32_t)seq;
        p.d = (uint32_t)d;
        p.dn = (uint32_t)((d + 15) / 16 * 16);
        p.num_kb_s = (uint32_t)(dpad / SBK);
        p.q_tiles = (uint32_t)((seq + BQ - 1) / BQ);
        p.scale = scale;
        p.causal = causal ? 1u : 0u;
        CUtensorMap mq_h, mq_l, mk_h, mk_l, mv_h, mv_l;
        TRN_TRY(make_map(&mq_h, q_hi, heads, seq, dpad, BQ, SBK));
        TRN_TRY(make_map(&mq_l, q_lo, heads, seq, dpad, BQ, SBK));
        // CTA pairs (two query tiles share every K / V tile) whenever a head has at least two query tiles.  With a causal
        // mask the pair also runs the earlier tile through the later tile's last key tile (wholly masked for it): 1/(T+1)
        // extra work at T query tiles against the pair kernel's ~18 % gain, so short causal sequences stay single.
        // TRN_ATT_PAIR=0 forces the single-CTA kernel (A/B measurements and its own parity tests), 2 forces pairs.
        static const int use_pair = [] { const char* e = getenv("TRN_ATT_PAIR"); return e ? atoi(e) : 1; }();
        const size_t q_pairs = (p.q_tiles + 1) / 2;
        const bool pair_pays = use_pair == 2 || !causal || p.q_tiles >= 6;
        if (use_pair && pair_pays && p.q_tiles >= 2 && heads * q_pairs * 2 <= 0x7FFFFFFFull) {
            TRN_TRY(make_map(&mk_h, k_hi, heads, seq, dpad, BKV / 2, SBK));
            TRN_TRY(make_map(&mk_l, k_lo, heads, seq, dpad, BKV / 2, SBK));
            TRN_TRY(make_map(&mv_h, v_hi, heads, d, seqpad, p.dn / 2, SBK));
            TRN_TRY(make_map(&mv_l, v_lo, heads, d, seqpad, p.dn / 2, SBK));
            static const cudaError_t optin2 = cudaFuncSetAttribute(attention_tf32x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
            TRN_CUDA(optin2);
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(heads * q_pairs * 2));
            cfg.blockDim = dim3(kPairThreads);
            cfg.dynamicSmemBytes = pair_smem_bytes(p.num_kb_s);
            cfg.stream = s;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2;
            attr[0].val.clusterDim.y = 1;
            attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr;
            cfg.numAttrs = 1;
            TRN_CUDA(cudaLaunchKernelEx(&cfg, attention_tf32x3_pair_kernel, mq_h, mq_l, mk_h, mk_l, mv_h, mv_l, p));
        } else {
            TRN_TRY(make_map(&mk_h, k_hi, heads, seq, dpad, BKV, SBK));
            TRN_TRY(make_map(&mk_l, k_lo, heads, seq, dpad, BKV, SBK));
            TRN_TRY(make_map(&mv_h, v_hi, heads, d, seqpad, p.dn, SBK));
            TRN_TRY(make_map(&mv_l, v_lo, heads, d, seqpad, p.dn, SBK));
            static const cudaError_t optin = cudaFuncSetAttribute(attention_tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBudget);
            TRN_CUDA(optin);
            attention_tf32x3_kernel<<<(unsigned)(heads * p.q_tiles), kThreads, smem_bytes(p.num_kb_s), s>>>(mq_h, mq_l, mk_h, mk_l, mv_h, mv_l, p);
        }
        count_launch();
        TRN_CUDA(cudaGetLastError());
        // IEEE path for Inf/NaN inputs: runs only when the split pre-pass raised the flag (checked on the device).  A
        // grid of a few blocks per SM striding over the rows: as a no-op it costs ~3 us (131 072 empty blocks cost 76 us)
        const size_t gated_blocks = std::min(simt_blocks, (size_t)cx->sm_count * 16);
        attention_simt_kernel<<<(unsigned)gated_blocks, 128, 0, s>>>(q, k, v, out, (uint32_t)heads, (uint32_t)seq, (uint32_t)d,
                                                                   scale, causal ? 1u : 0u, flag);
        count_launch();
        TRN_CUDA(cudaGetLastError());
        return TRN_OK;
    };
    st = run();
    scratch_free(scratch, s);
    return st;
}

}  // namespace trn
