// gemm_tc.cu — f32 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), 3xTF32.
//
// Replaces Matrix::matmul's blocked AVX2 path (matmul_simd + matmul_microkernel_4x1_avx2,
// src/matrix.rs:912-1401, :615-688) and the wgpu MATMUL_SHADER (src/backends/gpu/shaders.rs:11-49)
// for shapes that fill a tensor-core tile; batched_matmul / batched_matmul_4d (src/matrix.rs:383,
// :464) are the same kernel with a batch coordinate in the tile scheduler.
//
// Numerics (3xTF32): every f32 operand x is split once, in a streaming pre-pass, into
//   hi = tf32_round(x),  lo = tf32_round(x - hi)        (x - hi is exact in f32)
// and the product is accumulated as  lo_a*hi_b + hi_a*lo_b + hi_a*hi_b  in an f32 TMEM
// accumulator — three kind::tf32 MMAs per logical product.  The dropped lo*lo term is <= 2^-22
// relative per product.
//
// Two-level accumulation: the tensor core adds into its f32 accumulator with truncation, which
// is a BIASED rounding — measured on B200 the relative error of an all-positive product grows
// linearly with K (2.5e-5 at K=2048, 9.3e-5 at K=8192) when a whole tile is accumulated in TMEM.
// So TMEM only ever holds a partial sum over kChunkKB k-blocks (K = 128): the MMA warp alternates
// between the two accumulator stages, and the epilogue warps drain each finished partial into f32
// REGISTER accumulators with round-to-nearest adds while the next partial is being computed.  The
// bias is then bounded by the chunk length (48 accumulate steps, ~1.5e-6 relative) instead of K.
//
// Non-finite inputs: hi*lo would manufacture Inf*0 = NaN where IEEE gives Inf, so the split
// pre-pass raises a device flag when it meets an Inf/NaN; the tensor-core kernel then exits at
// once and the SIMT FFMA kernel (exact IEEE semantics) computes the product instead.  Both
// kernels are always enqueued, the flag is read on the device: no host round trip.
// The same pre-pass re-lays B (k x n, row-major) out K-major (n x k), so both operands use the
// canonical K-major SWIZZLE_128B shared-memory layout, and pads K to a multiple of 32 with
// zeros so the main loop has no remainder and TMA strides are 16-byte aligned.
//
// Kernel: persistent, warp-specialised, one CTA per SM.
//   warp 0 : TMA producer   (cp.async.bulk.tensor.3d -> smem ring, mbarrier complete_tx)
//   warp 1 : MMA issuer     (one elected lane; tcgen05.mma.cta_group::1.kind::tf32, 128x256x8),
//            also owns the TMEM allocation (512 columns = two 128x256 f32 accumulators)
//   warps 2-9 : epilogue    (tcgen05.ld 32x32b -> f32 register accumulators -> swizzled smem block ->
//            cp.async.bulk.tensor store); warp w owns TMEM lane quadrant w%4 and column half (w-2)/4
//            of the 128x256 tile.
// Accumulation order is fixed (k ascending, no split-K, no atomics) => bit-identical reruns
// (tests/wasm_optimization_tests.rs:200-230).
//
// Roofline: tensor pipe.  Algorithmic work 2*m*n*k flop; the pipe executes 3x that in TF32.
#include <cuda.h>

#include <cstdlib>

#include <algorithm>
#include <cstdio>
#include <vector>

#include "common.cuh"
#include "tcgen05.cuh"

namespace trn {
namespace tc {

constexpr int BM = 128;   // UMMA M (cta_group::1)
constexpr int BN = 256;   // UMMA N
constexpr int BK = 32;    // K padding granularity of the split operands (Kpad % 32 == 0)
constexpr int kThreads = 320;  // TMA warp + MMA warp + 8 epilogue warps
constexpr int kPairThreads = 384;  // pair kernel: warpgroup 0 = TMA, MMA, 2 idle warps; warpgroups 1-2 = 8 epilogue warps
constexpr int kEpiWarps = 8;
constexpr uint32_t kChunkK = 128;  // K extent accumulated in TMEM before draining to registers
constexpr uint32_t kTmemCols = 512;        // 2 accumulator stages x 256 columns

// SBK = floats of K per pipeline stage: 32 (128-byte rows, SWIZZLE_128B) or 16 (64-byte rows, SWIZZLE_64B).
// 3xTF32 moves hi+lo of both operands per stage, so SBK = 16 buys a 4-deep ring (4 x 48 KiB) where SBK = 32
// only fits 2 x 96 KiB: with 3 stages in flight instead of 1 the ring tolerates ~2300 cycles of load latency.
template <int TERMS, int SBK>
struct Cfg {
    static constexpr int kOperands = TERMS == 3 ? 2 : 1;  // hi (+ lo)
    static constexpr uint32_t kABytes = BM * SBK * 4;
    static constexpr uint32_t kBBytes = BN * SBK * 4;
    static constexpr uint32_t kStageBytes = kOperands * (kABytes + kBBytes);
    static constexpr int kStages = (TERMS == 3 ? 2 : 4) * (32 / SBK);
    static constexpr uint32_t kChunkKB = kChunkK / SBK;       // k-blocks per TMEM partial sum
    static constexpr uint32_t kStoreBytes = kEpiWarps * 4096;  // one 32x32 f32 staging block per epilogue warp
    static constexpr uint32_t kSmemBytes = kStages * kStageBytes + kStoreBytes + 1024 /*align*/ + 256 /*barriers*/;
};

struct Params {
    float* c;
    uint32_t tma_store;         // 1: C goes out through smem + cp.async.bulk.tensor stores (needs n % 4 == 0)
    const int* nonfinite_flag;  // set by the split pre-pass; non-zero => this kernel must not run
    uint32_t m, n;            // C rows / cols per batch
    uint32_t num_kb;          // Kpad / BK
    uint32_t tiles_m, tiles_n, batch;
    uint32_t nseg;            // A-stationary kernel: n-segments per A panel (work unit = panel x segment)
    uint32_t rot_mult;        // A-stationary kernel: pair p starts its units at n-tile (p * rot_mult) mod (tiles of the unit)
    uint32_t store_hint;      // fused kernels: 1 = C stores carry an L2 evict_first policy (C much larger than L2)
    uint32_t debug;           // fused kernels, experiments (TRN_GEMM_DEBUG): bit 0 skip the C stores, bit 1 aim every store at batch 0
    uint32_t terms_mask;      // fused kernel, debugging: bit 0 lo*hi, bit 1 hi*lo, bit 2 hi*hi (7 = the product)
    unsigned long long* trace;  // A-stationary kernel, experiments (TRN_GEMM_TRACE): per-CTA %globaltimer at entry / exit
};

// Tile order: batch-major, then groups of 16 m-tiles, n fastest-but-one inside a group, so CTAs
// that run concurrently share A row-panels and B column-panels in L2.
template <uint32_t G = 16>
__device__ __forceinline__ void tile_coords(uint32_t t, const Params& p, uint32_t& b, uint32_t& mt, uint32_t& nt) {
    const uint32_t per_batch = p.tiles_m * p.tiles_n;
    b = t / per_batch;
    uint32_t r = t % per_batch;
    const uint32_t group = r / (G * p.tiles_n);
    const uint32_t first_m = group * G;
    const uint32_t gsize = min(G, p.tiles_m - first_m);
    const uint32_t in_group = r - group * G * p.tiles_n;
    mt = first_m + in_group % gsize;
    nt = in_group / gsize;
}

template <int TERMS, int SBK>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tf32_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                 const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                 const __grid_constant__ CUtensorMap map_c, const Params p) {
    using C = Cfg<TERMS, SBK>;
    constexpr uint32_t kABytes = C::kABytes, kBBytes = C::kBBytes, kChunkKB = C::kChunkKB;
    if (*p.nonfinite_flag != 0) return;  // Inf/NaN in the inputs: the SIMT kernel takes over (grid-uniform)
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles must sit on 1024-byte boundaries
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t store_base = smem_base + C::kStages * C::kStageBytes;   // 1024-aligned staging for C
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_gen + C::kStages * C::kStageBytes + C::kStoreBytes);
    // barrier slots: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2], then the TMEM base
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (C::kStages + s); };
    auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + s); };
    auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * C::kStages + 2 + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t total_tiles = p.tiles_m * p.tiles_n * p.batch;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi);
        tma_prefetch_desc(&map_b_hi);
        if (TERMS == 3) { tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_b_lo); }
        for (int s = 0; s < C::kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tmem_full_bar(s), 1); mbar_init(tmem_empty_bar(s), kEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(smem_u32(tmem_slot), kTmemCols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
                uint32_t b, mt, nt;
                tile_coords(t, p, b, mt, nt);
                const int m0 = (int)(mt * BM), n0 = (int)(nt * BN);
                for (uint32_t kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * C::kStageBytes;
                    mbar_expect_tx(full_bar(stage), C::kStageBytes);
                    const int k0 = (int)(kb * SBK);
                    tma_load_3d(sa, &map_a_hi, full_bar(stage), k0, m0, (int)b);
                    tma_load_3d(sa + kABytes, &map_b_hi, full_bar(stage), k0, n0, (int)b);
                    if (TERMS == 3) {
                        tma_load_3d(sa + kABytes + kBBytes, &map_a_lo, full_bar(stage), k0, m0, (int)b);
                        tma_load_3d(sa + 2 * kABytes + kBBytes, &map_b_lo, full_bar(stage), k0, n0, (int)b);
                    }
                    if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = make_idesc_tf32(BM, BN);
        uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
        for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            for (uint32_t kb = 0; kb < p.num_kb; ++kb) {
                const uint32_t in_chunk = kb % kChunkKB;
                if (in_chunk == 0) {
                    mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);  // epilogue has drained this accumulator
                    tc_fence_after();
                }
                const uint32_t d = tmem_base + acc * BN;
                mbar_wait(full_bar(stage), phase);
                tc_fence_after();
                const bool chunk_end = in_chunk == kChunkKB - 1 || kb == p.num_kb - 1;
                if (elect_one()) {
                    const uint32_t sa = smem_base + stage * C::kStageBytes;
                    const uint32_t a_hi = sa, b_hi = sa + kABytes;
                    const uint32_t a_lo = sa + kABytes + kBBytes, b_lo = sa + 2 * kABytes + kBBytes;
#pragma unroll
                    for (int k = 0; k < SBK / UMMA_K; ++k) {
                        const uint32_t koff = k * UMMA_K * 4;  // 32 bytes along K inside the swizzle span
                        const uint32_t accum = (in_chunk | (uint32_t)k) != 0;  // first MMA of a chunk overwrites
                        if (TERMS == 3) {
                            umma_tf32(d, make_desc_k<SBK>(a_lo + koff), make_desc_k<SBK>(b_hi + koff), idesc, accum);
                            umma_tf32(d, make_desc_k<SBK>(a_hi + koff), make_desc_k<SBK>(b_lo + koff), idesc, 1u);
                            umma_tf32(d, make_desc_k<SBK>(a_hi + koff), make_desc_k<SBK>(b_hi + koff), idesc, 1u);
                        } else {
                            umma_tf32(d, make_desc_k<SBK>(a_hi + koff), make_desc_k<SBK>(b_hi + koff), idesc, accum);
                        }
                    }
                    umma_commit(empty_bar(stage));                 // smem slot free once these MMAs retire
                    if (chunk_end) umma_commit(tmem_full_bar(acc));  // partial sum complete
                }
                __syncwarp();
                if (++stage == C::kStages) { stage = 0; phase ^= 1; }
                if (chunk_end && ++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else {
        // ===================== epilogue (warps 2..9) =====================
        const uint32_t quad = warp & 3;          // TMEM lane quadrant this warp may access
        const uint32_t half = (warp - 2) >> 2;   // which 128 of the 256 accumulator columns
        const uint32_t num_chunks = (p.num_kb + kChunkKB - 1) / kChunkKB;
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t t = blockIdx.x; t < total_tiles; t += gridDim.x) {
            uint32_t b, mt, nt;
            tile_coords(t, p, b, mt, nt);
            float sum[128];  // this thread's row, 128 consecutive columns: the second accumulation level
            for (uint32_t ch = 0; ch < num_chunks; ++ch) {
                mbar_wait(tmem_full_bar(acc), acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((quad * 32u) << 16) + acc * BN + half * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + j * 32, r);
                    tmem_ld_wait();
                    if (ch == 0) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) sum[j * 32 + i] = __uint_as_float(r[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) sum[j * 32 + i] = __fadd_rn(sum[j * 32 + i], __uint_as_float(r[i]));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tmem_empty_bar(acc));
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            // ---- store this thread's 128 columns
            const uint32_t row0 = mt * BM + quad * 32;
            const uint32_t row = row0 + lane;
            const uint32_t col_base = nt * BN + half * 128;
            if (p.tma_store) {
                // registers -> 32x32 block in smem (SWIZZLE_128B pattern: 16-byte chunk j of row r sits at
                // chunk j ^ (r & 7), conflict-free for row-per-lane float4 writes) -> TMA tensor store, which
                // writes full 128-byte lines and clips rows >= m / cols >= n by itself.
                const uint32_t stage_addr = store_base + (uint32_t)(warp - 2) * 4096u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // staging free again
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint32_t dst = stage_addr + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) * 16);
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(dst), "f"(sum[j * 32 + 4 * q]),
                                     "f"(sum[j * 32 + 4 * q + 1]), "f"(sum[j * 32 + 4 * q + 2]), "f"(sum[j * 32 + 4 * q + 3])
                                     : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0 && row0 < p.m && col_base + j * 32 < p.n) {
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                                     :: "l"(&map_c), "r"(stage_addr), "r"((int)(col_base + j * 32)), "r"((int)row0), "r"((int)b)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            } else if (row < p.m && col_base < p.n) {
                float* crow = p.c + ((size_t)b * p.m + row) * p.n + col_base;
#pragma unroll
                for (int j = 0; j < 128; ++j)
                    if (col_base + j < p.n) crow[j] = sum[j];
            }
        }
        // all of this warp's tensor stores must have landed before the CTA exits
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// =====================================================================================================
// CTA-pair kernel: tcgen05.mma.cta_group::2 — two SMs of one TPC compute one 256 x 256 tile.
//
// Each CTA of the 2-CTA cluster loads ITS 128 rows of A and ITS 128 of the tile's 256 B rows (n), i.e.
// 32 KiB per 16-wide k-block instead of 48 KiB: one third less L2->SM traffic and a 6-deep ring in the
// same shared memory (5 stages = 3840 MMA-cycles of loads in flight).  The leader CTA (cluster rank 0)
// issues every MMA (M = 256, N = 256, K = 8; A/B descriptors name the leader's shared memory, the tensor
// core reads the peer's at the same offsets) and the accumulator rows land in each CTA's own TMEM, so
// both CTAs run the same two-level epilogue on their own 128 rows.
// Synchronisation: TMA loads of both CTAs complete_tx on the LEADER's full barrier; tcgen05.commit
// multicasts the "stage free" / "accumulator full" arrivals to both CTAs; the epilogue warps of both CTAs
// arrive on the leader's "accumulator empty" barrier through the cluster shared window.
// =====================================================================================================
namespace pair {
constexpr int SBK = 16;
constexpr uint32_t kABytes = BM * SBK * 4;        // 8 KiB: this CTA's 128 rows of A
constexpr uint32_t kBBytes = (BN / 2) * SBK * 4;  // 8 KiB: this CTA's 128 of the 256 B rows
constexpr uint32_t kStageBytes = 2 * (kABytes + kBBytes);   // hi + lo: 32 KiB per CTA
constexpr int kStages = 6;
constexpr uint32_t kChunkKB = kChunkK / SBK;
constexpr uint32_t kStoreBytes = kEpiWarps * 4096;
constexpr uint32_t kSmemBytes = kStages * kStageBytes + kStoreBytes + 1024 + 256;
}  // namespace pair

__global__ void __launch_bounds__(kPairThreads, 1)
gemm_tf32x3_pair_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
                        const __grid_constant__ CUtensorMap map_b_hi, const __grid_constant__ CUtensorMap map_b_lo,
                        const __grid_constant__ CUtensorMap map_c, const Params p) {
    using namespace pair;
    if (*p.nonfinite_flag != 0) return;  // grid-uniform: both CTAs of every pair leave together
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t store_base = smem_base + kStages * kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_gen + kStages * kStageBytes + kStoreBytes);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto tmem_full_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
    auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (2 * kStages + 2 + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();          // 0 = leader
    const uint32_t pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const uint32_t total_tiles = p.tiles_m * p.tiles_n * p.batch;   // tiles_m counts 256-row pair tiles here

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a_hi); tma_prefetch_desc(&map_b_hi);
        tma_prefetch_desc(&map_a_lo); tma_prefetch_desc(&map_b_lo);
        for (int s = 0; s < kStages; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(tmem_full_bar(s), 1); mbar_init(tmem_empty_bar(s), 2 * kEpiWarps); }
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

    // Register budget by role (setmaxnreg is per warpgroup, and ptxas honours it only as the first statement of the role's
    // branch): warpgroup 0 = TMA warp, MMA warp, two idle warps; warpgroups 1-2 = the epilogue warps, whose 128 register
    // accumulators of the second accumulation level do not fit the uniform 168-register limit.
    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;" ::: "memory");
    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = pair_id; t < total_tiles; t += num_pairs) {
                uint32_t b, mt, nt;
                tile_coords<8>(t, p, b, mt, nt);
                const int m0 = (int)(mt * 2 * BM + rank * BM), n0 = (int)(nt * BN + rank * (BN / 2));
                for (uint32_t kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * kStageBytes;
                    const uint32_t lead_full = full_bar(stage) & kPeerMask;
                    if (rank == 0) mbar_expect_tx(full_bar(stage), 2 * kStageBytes);   // both CTAs' bytes
                    const int k0 = (int)(kb * SBK);
                    tma_load_3d_pair(sa, &map_a_hi, lead_full, k0, m0, (int)b);
                    tma_load_3d_pair(sa + kABytes, &map_b_hi, lead_full, k0, n0, (int)b);
                    tma_load_3d_pair(sa + kABytes + kBBytes, &map_a_lo, lead_full, k0, m0, (int)b);
                    tma_load_3d_pair(sa + 2 * kABytes + kBBytes, &map_b_lo, lead_full, k0, n0, (int)b);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(2 * BM, BN);
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (uint32_t t = pair_id; t < total_tiles; t += num_pairs) {
                for (uint32_t kb = 0; kb < p.num_kb; ++kb) {
                    const uint32_t in_chunk = kb % kChunkKB;
                    if (in_chunk == 0) {
                        mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);   // both CTAs' epilogues have drained it
                        tc_fence_after();
                    }
                    const uint32_t d = tmem_base + acc * BN;
                    mbar_wait(full_bar(stage), phase);
                    tc_fence_after();
                    const bool chunk_end = in_chunk == kChunkKB - 1 || kb == p.num_kb - 1;
                    if (elect_one()) {
                        const uint32_t sa = smem_base + stage * kStageBytes;
                        const uint32_t a_hi = sa, b_hi = sa + kABytes;
                        const uint32_t a_lo = sa + kABytes + kBBytes, b_lo = sa + 2 * kABytes + kBBytes;
#pragma unroll
                        for (int k = 0; k < SBK / UMMA_K; ++k) {
                            const uint32_t koff = k * UMMA_K * 4;
                            const uint32_t accum = (in_chunk | (uint32_t)k) != 0;
                            umma_tf32_pair(d, make_desc_k<SBK>(a_lo + koff), make_desc_k<SBK>(b_hi + koff), idesc, accum);
                            umma_tf32_pair(d, make_desc_k<SBK>(a_hi + koff), make_desc_k<SBK>(b_lo + koff), idesc, 1u);
                            umma_tf32_pair(d, make_desc_k<SBK>(a_hi + koff), make_desc_k<SBK>(b_hi + koff), idesc, 1u);
                        }
                        umma_commit_pair(empty_bar(stage));
                        if (chunk_end) umma_commit_pair(tmem_full_bar(acc));
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                    if (chunk_end && ++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    }
    } else {
        // ===================== epilogue (warps 4..11, both CTAs, own 128 rows) =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;" ::: "memory");
        const uint32_t quad = warp & 3;
        const uint32_t half = (warp - 4) >> 2;
        const uint32_t num_chunks = (p.num_kb + kChunkKB - 1) / kChunkKB;
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t t = pair_id; t < total_tiles; t += num_pairs) {
            uint32_t b, mt, nt;
            tile_coords<8>(t, p, b, mt, nt);
            float sum[128];
            for (uint32_t ch = 0; ch < num_chunks; ++ch) {
                mbar_wait(tmem_full_bar(acc), acc_phase);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((quad * 32u) << 16) + acc * BN + half * 128;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + j * 32, r);
                    tmem_ld_wait();
                    if (ch == 0) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) sum[j * 32 + i] = __uint_as_float(r[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) sum[j * 32 + i] = __fadd_rn(sum[j * 32 + i], __uint_as_float(r[i]));
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(tmem_empty_bar(acc) & kPeerMask);   // on the leader's barrier
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
            }
            const uint32_t row0 = mt * 2 * BM + rank * BM + quad * 32;
            const uint32_t row = row0 + lane;
            const uint32_t col_base = nt * BN + half * 128;
            if (p.tma_store) {
                const uint32_t stage_addr = store_base + (uint32_t)(warp - 4) * 4096u;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint32_t dst = stage_addr + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) * 16);
                        asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(dst), "f"(sum[j * 32 + 4 * q]),
                                     "f"(sum[j * 32 + 4 * q + 1]), "f"(sum[j * 32 + 4 * q + 2]), "f"(sum[j * 32 + 4 * q + 3])
                                     : "memory");
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0 && row0 < p.m && col_base + j * 32 < p.n) {
                        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                                     :: "l"(&map_c), "r"(stage_addr), "r"((int)(col_base + j * 32)), "r"((int)row0), "r"((int)b)
                                     : "memory");
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                }
            } else if (row < p.m && col_base < p.n) {
                float* crow = p.c + ((size_t)b * p.m + row) * p.n + col_base;
#pragma unroll
                for (int j = 0; j < 128; ++j)
                    if (col_base + j < p.n) crow[j] = sum[j];
            }
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
    }

    // neither CTA may leave (or free TMEM) while its partner can still signal it or read its shared memory
    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// =====================================================================================================
// Fused-split CTA-pair kernel: the same cta_group::2 tile machine WITHOUT the operand pre-pass.
//
// Why: at short K the pre-pass is a large share of the call (config 3, K = 128: 0.24 of 1.6 ms) and, worse, the
// pre-split operands double the L2 -> SM traffic of every tile.  Config 3 moves 256 KiB of split operands in and
// 128 KiB of C out per tile and SM, 62 B/clk/SM against an L2 slice throughput of ~42 B/clk/SM (6300 B/clk per chip):
// the tile period was 9000 cycles for 6144 cycles of math.  Here TMA brings the RAW f32 tiles (half the bytes):
//   * hi operand = the raw f32 words as they lie: kind::tf32 reads the top 19 bits of each word, i.e. the tensor
//     core itself truncates x to hi = x & 0xFFFFE000;
//   * lo operand = tf32_rna(x - hi), computed by two splitter warps per CTA from shared memory into a second copy
//     of the tile AT THE SAME OFFSETS — an element-wise pass over the swizzled bytes that needs no knowledge of the
//     layout.  x - hi is exact (13 bits) and its rounding to 11 bits costs 2^-10 * 2^-12 = 2^-22 of |x|, the same
//     error as the pre-pass's round-to-nearest hi / lo pair; the dropped lo*lo term is <= 2^-20 relative.
//   * A stays K-major ([m][k] rows, SWIZZLE_64B boxes of 128 x 16); B is consumed as it lies in memory, [k][n]
//     row-major = MN-major for the tensor core (instruction-descriptor bit 16).  For 32-bit operands the only MN-major
//     shared-memory layout is the 128-byte swizzle with 32-byte atoms (descriptor layout type 1, TMA swizzle
//     128B_ATOM_32B: atoms of 32 n x 4 k; TMA boxes of 32 n x 16 k), so there is no transpose either.
// K, m and n tails are zero-filled by TMA (no padding pass); Inf/NaN inputs are detected by the splitter warps,
// which raise the same device flag the pre-pass raises: the gated SIMT kernel enqueued behind this one then
// recomputes C with IEEE semantics.  Needs k % 4 == 0, n % 4 == 0 and 16-byte aligned operands (TMA strides).
// Pipeline per stage: TMA (full) -> splitter warps of BOTH CTAs (ready, on the leader) -> MMA -> commit (empty).
// =====================================================================================================
// One tile of the CTA-pair epilogue (shared by the fused-split kernels): drain this CTA's 128 rows of the accumulator,
// chunk by chunk, into register accumulators (the second accumulation level), then registers -> swizzled 32x32 staging
// block -> TMA tensor store.  `ewarp` 0..7: TMEM lane quadrant ewarp % 4, column half ewarp / 4.
__device__ __forceinline__ void pair_epilogue_tile(const CUtensorMap* map_c, const Params& p, uint32_t tmem_base, uint32_t store_base,
                                                   uint32_t tmem_full0, uint32_t tmem_empty0, uint32_t num_chunks, uint32_t& acc,
                                                   uint32_t& acc_phase, uint32_t b, uint32_t mt, uint32_t nt, int ewarp, int lane,
                                                   uint32_t rank) {
    const uint32_t quad = (uint32_t)ewarp & 3u, half = (uint32_t)ewarp >> 2;
    auto cblock = [&](int j) -> uint32_t { return half * 4u + (uint32_t)j; };   // the 32-column blocks this warp owns
    float sum[128];
    for (uint32_t ch = 0; ch < num_chunks; ++ch) {
        mbar_wait(tmem_full0 + 8u * acc, acc_phase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((quad * 32u) << 16) + acc * BN;
        // (one load, one wait, four times: issuing the four loads of a first partial back to back and waiting once measured
        // 1.4 % SLOWER on config 3 — 1.271 vs 1.254 ms, three alternating runs, scripts/exp/exp_cfg3_ab.py)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            uint32_t r[32];
            tmem_ld_32x32(taddr + cblock(j) * 32, r);
            tmem_ld_wait();
            if (ch == 0) {
#pragma unroll
                for (int i = 0; i < 32; ++i) sum[j * 32 + i] = __uint_as_float(r[i]);
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) sum[j * 32 + i] = __fadd_rn(sum[j * 32 + i], __uint_as_float(r[i]));
            }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster((tmem_empty0 + 8u * acc) & kPeerMask);   // on the leader's barrier
        if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    const uint32_t row0 = mt * 2 * BM + rank * BM + quad * 32;
    const uint32_t row = row0 + lane;
    const uint32_t col_base = nt * BN + half * 128;
    if (p.tma_store) {
        const uint32_t stage_addr = store_base + (uint32_t)ewarp * 4096u;
        // C larger than L2 is a pure stream: evict_first keeps it from displacing the operand tiles and lets L2 write it
        // back in arrival order (config 3, 4 GiB of C: 1.325 -> 1.27 ms; no effect in the pre-split kernel of round 1)
        uint64_t policy = 0;
        if (p.store_hint) policy = l2_policy_drop();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint32_t dst = stage_addr + (uint32_t)lane * 128u + (uint32_t)((q ^ (lane & 7)) * 16);
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" :: "r"(dst), "f"(sum[j * 32 + 4 * q]),
                             "f"(sum[j * 32 + 4 * q + 1]), "f"(sum[j * 32 + 4 * q + 2]), "f"(sum[j * 32 + 4 * q + 3])
                             : "memory");
            }
            const uint32_t col = nt * BN + cblock(j) * 32;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0 && row0 < p.m && col < p.n && !(p.debug & 1u)) {
                if (p.store_hint)
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4}], [%1], %5;"
                                 :: "l"(map_c), "r"(stage_addr), "r"((int)col), "r"((int)row0), "r"((int)b), "l"(policy) : "memory");
                else
                asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                             :: "l"(map_c), "r"(stage_addr), "r"((int)col), "r"((int)row0), "r"((int)((p.debug & 2u) ? 0u : b))
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    } else if (row < p.m && col_base < p.n) {
        float* crow = p.c + ((size_t)b * p.m + row) * p.n + col_base;
#pragma unroll
        for (int j = 0; j < 128; ++j)
            if (col_base + j < p.n) crow[j] = sum[j];
    }
}

namespace fused {
constexpr int SBK = 16;
constexpr uint32_t kARaw = BM * SBK * 4;            // 8 KiB: this CTA's 128 rows of A, K-major
constexpr uint32_t kBBox = 32 * SBK * 4;            // 2 KiB: one TMA box of B, 32 n x 16 k (four 512-byte swizzle atoms)
constexpr uint32_t kBRaw = (BN / 2 / 32) * kBBox;   // 8 KiB: this CTA's 128 of the tile's 256 columns of B
constexpr uint32_t kRawBytes = kARaw + kBRaw;       // what TMA lands per stage
constexpr uint32_t kStageBytes = 2 * kRawBytes;     // raw + lo
constexpr int kStages = 6;
constexpr uint32_t kChunkKB = kChunkK / SBK;
constexpr uint32_t kStoreBytes = kEpiWarps * 4096;
constexpr uint32_t kSmemBytes = kStages * kStageBytes + kStoreBytes + 1024 + 256;
constexpr int kSplitWarps = 2;                      // warps 2-3 of warpgroup 0
}  // namespace fused

__device__ __forceinline__ uint32_t split_lo_bits(uint32_t x, bool& bad) {
    bad |= (x & 0x7F800000u) == 0x7F800000u;
    const float r = __uint_as_float(x) - __uint_as_float(x & 0xFFFFE000u);   // exact
    uint32_t l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
    return l;
}

__global__ void __launch_bounds__(kPairThreads, 1)
gemm_tf32x3_fused_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                              const __grid_constant__ CUtensorMap map_c, const Params p, int* raise_flag) {
    using namespace fused;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t store_base = smem_base + kStages * kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_gen + kStages * kStageBytes + kStoreBytes);
    const uint32_t bar_base = smem_u32(bars);
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (kStages + s); };
    auto ready_bar = [&](int s) { return bar_base + 8u * (2 * kStages + s); };
    auto tmem_full_bar = [&](int s) { return bar_base + 8u * (3 * kStages + s); };
    auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (3 * kStages + 2 + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    const uint32_t total_tiles = p.tiles_m * p.tiles_n * p.batch;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b);
        for (int s = 0; s < kStages; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
            mbar_init(ready_bar(s), 2 * kSplitWarps);   // the splitter warps of both CTAs (leader's copy is the one used)
        }
        for (int s = 0; s < 2; ++s) { mbar_init(tmem_full_bar(s), 1); mbar_init(tmem_empty_bar(s), 2 * kEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;" ::: "memory");
    if (warp == 0) {
        // ===================== TMA producer (both CTAs, own tiles, own barrier) =====================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            for (uint32_t t = pair_id; t < total_tiles; t += num_pairs) {
                uint32_t b, mt, nt;
                tile_coords<8>(t, p, b, mt, nt);
                const int m0 = (int)(mt * 2 * BM + rank * BM), n0 = (int)(nt * BN + rank * (BN / 2));
                for (uint32_t kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(empty_bar(stage), phase ^ 1);
                    const uint32_t sa = smem_base + stage * kStageBytes;
                    mbar_expect_tx(full_bar(stage), kRawBytes);
                    const int k0 = (int)(kb * SBK);
                    tma_load_3d(sa, &map_a, full_bar(stage), k0, m0, (int)b);
#pragma unroll
                    for (int j = 0; j < BN / 2 / 32; ++j)
                        tma_load_3d(sa + kARaw + j * kBBox, &map_b, full_bar(stage), n0 + 32 * j, k0, (int)b);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(2 * BM, BN) | (1u << 16);   // B is MN-major
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
            for (uint32_t t = pair_id; t < total_tiles; t += num_pairs) {
                for (uint32_t kb = 0; kb < p.num_kb; ++kb) {
                    const uint32_t in_chunk = kb % kChunkKB;
                    if (in_chunk == 0) {
                        mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);
                        tc_fence_after();
                    }
                    const uint32_t d = tmem_base + acc * BN;
                    mbar_wait(ready_bar(stage), phase);   // both CTAs' stages are landed and split
                    tc_fence_after();
                    const bool chunk_end = in_chunk == kChunkKB - 1 || kb == p.num_kb - 1;
                    if (elect_one()) {
                        const uint32_t sa = smem_base + stage * kStageBytes;
                        const uint32_t a_hi = sa, b_hi = sa + kARaw;
                        const uint32_t a_lo = sa + kRawBytes, b_lo = a_lo + kARaw;
#pragma unroll
                        for (int k = 0; k < SBK / UMMA_K; ++k) {
                            const uint32_t ka = k * UMMA_K * 4;   // A: 32 bytes along K inside the 64-byte swizzled row
                            const uint32_t kbo = k * 1024;        // B: the next 8 k-rows (two 4-row swizzle atoms) of every box
                            const uint32_t accum = (in_chunk | (uint32_t)k) != 0;
                            if (p.terms_mask == 7u) {
                                umma_tf32_pair(d, make_desc_k<SBK>(a_lo + ka), make_desc_mn_tf32(b_hi + kbo, kBBox), idesc, accum);
                                umma_tf32_pair(d, make_desc_k<SBK>(a_hi + ka), make_desc_mn_tf32(b_lo + kbo, kBBox), idesc, 1u);
                                umma_tf32_pair(d, make_desc_k<SBK>(a_hi + ka), make_desc_mn_tf32(b_hi + kbo, kBBox), idesc, 1u);
                            } else {   // TRN_GEMM_FUSED_TERMS: single terms for debugging / the 1xTF32 probe
                                uint32_t acc1 = accum;
                                if (p.terms_mask & 1u) { umma_tf32_pair(d, make_desc_k<SBK>(a_lo + ka), make_desc_mn_tf32(b_hi + kbo, kBBox), idesc, acc1); acc1 = 1u; }
                                if (p.terms_mask & 2u) { umma_tf32_pair(d, make_desc_k<SBK>(a_hi + ka), make_desc_mn_tf32(b_lo + kbo, kBBox), idesc, acc1); acc1 = 1u; }
                                if (p.terms_mask & 4u) { umma_tf32_pair(d, make_desc_k<SBK>(a_hi + ka), make_desc_mn_tf32(b_hi + kbo, kBBox), idesc, acc1); }
                            }
                        }
                        umma_commit_pair(empty_bar(stage));
                        if (chunk_end) umma_commit_pair(tmem_full_bar(acc));
                    }
                    __syncwarp();
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                    if (chunk_end && ++acc == 2) { acc = 0; acc_phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== splitter warps (2-3, both CTAs): lo = tf32_rna(x - trunc_tf32(x)) =====================
        const uint32_t sw = (uint32_t)warp - 2;
        bool bad = false;
        uint32_t stage = 0, phase = 0;
        constexpr uint32_t kPerWarp = kRawBytes / 16 / kSplitWarps;   // float4 per warp per stage (512)
        for (uint32_t t = pair_id; t < total_tiles; t += num_pairs) {
            for (uint32_t kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(full_bar(stage), phase);
                const uint32_t raw = smem_base + stage * kStageBytes + (sw * kPerWarp + (uint32_t)lane) * 16u;
#pragma unroll
                for (uint32_t h = 0; h < kPerWarp / 32 / 8; ++h) {
                    uint32_t v[8][4];
#pragma unroll
                    for (int u = 0; u < 8; ++u)
                        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3])
                                     : "r"(raw + (h * 8 + u) * 512u));
#pragma unroll
                    for (int u = 0; u < 8; ++u) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) v[u][e] = split_lo_bits(v[u][e], bad);
                        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(raw + kRawBytes + (h * 8 + u) * 512u), "r"(v[u][0]),
                                     "r"(v[u][1]), "r"(v[u][2]), "r"(v[u][3]) : "memory");
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tensor core reads lo through the async proxy
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(ready_bar(stage) & kPeerMask);
                if (++stage == kStages) { stage = 0; phase ^= 1; }
            }
        }
        if (bad) *raise_flag = 1;
    }
    } else {
        // ===================== epilogue (warps 4..11, both CTAs, own 128 rows) =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;" ::: "memory");
        const uint32_t num_chunks = (p.num_kb + kChunkKB - 1) / kChunkKB;
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t t = pair_id; t < total_tiles; t += num_pairs) {
            uint32_t b, mt, nt;
            tile_coords<8>(t, p, b, mt, nt);
            pair_epilogue_tile(&map_c, p, tmem_base, store_base, tmem_full_bar(0), tmem_empty_bar(0), num_chunks, acc, acc_phase,
                               b, mt, nt, warp - 4, lane, rank);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// =====================================================================================================
// A-stationary fused-split kernel for K <= 128 (the batched Q K^T shape of config 3: 256 heads of 2048 x 128 x 2048).
//
// ncu on the fused kernel above at config 3: tile period 9700 cycles for 6144 cycles of math, the SM's shared-memory data
// pipe ~97 % busy — per tile and CTA 3077 wavefronts (128 B) of MMA operand reads, 3072 of splitter / staging LSU traffic,
// 2048 of TMA writes (operands) and reads (C stores), plus arbitration losses.  This kernel keeps the pair's 256-row A panel
// (raw + lo, 128 KiB per CTA at K = 128) in shared memory for ALL the n-tiles of its row of C and rings only B (4 stages
// of 16 KiB): per tile that removes 7/8 of A's TMA writes and of A's splitter reads and writes (1344 wavefronts) and
// halves the L2 -> SM operand traffic again.  A pair works through whole (panel, n-segment) units; the next panel is
// loaded k-block by k-block into the slots the LAST tile of the current unit frees (a_empty[kb], committed only in that
// tile), which hides the reload under that tile's remaining MMAs.
// =====================================================================================================
namespace astat {
constexpr int SBK = 16;
constexpr int kMaxKB = 8;                            // K <= 128
constexpr uint32_t kABlk = BM * SBK * 4;             // 8 KiB: one k-block of this CTA's 128 A rows (raw; the lo copy follows it)
constexpr uint32_t kASlot = 2 * kABlk;               // 16 KiB per k-block: raw + lo
constexpr uint32_t kPanelBytes = kMaxKB * kASlot;    // 128 KiB
constexpr uint32_t kBBox = 32 * SBK * 4;             // 2 KiB: 32 n x 16 k
constexpr uint32_t kBRaw = (BN / 2 / 32) * kBBox;    // 8 KiB: this CTA's 128 columns of one k-block of B
constexpr uint32_t kBStage = 2 * kBRaw;              // raw + lo
constexpr int kBStages = 4;
constexpr uint32_t kStoreBytes = kEpiWarps * 4096;
constexpr uint32_t kSmemBytes = kPanelBytes + kBStages * kBStage + kStoreBytes + 1024 + 512;
constexpr int kSplitWarps = 2;
}  // namespace astat

// one 8 KiB block (512 float4) over the two splitter warps: 8 float4 per lane, all loads in flight before the first store
__device__ __forceinline__ void split_block_8k(uint32_t raw_addr, uint32_t lo_addr, uint32_t sw, int lane, bool& bad) {
    const uint32_t off = (sw * 256u + (uint32_t)lane) * 16u;
    uint32_t v[8][4];
#pragma unroll
    for (int u = 0; u < 8; ++u)
        asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u][0]), "=r"(v[u][1]), "=r"(v[u][2]), "=r"(v[u][3])
                     : "r"(raw_addr + off + u * 512u));
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
        for (int e = 0; e < 4; ++e) v[u][e] = split_lo_bits(v[u][e], bad);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" :: "r"(lo_addr + off + u * 512u), "r"(v[u][0]), "r"(v[u][1]),
                     "r"(v[u][2]), "r"(v[u][3]) : "memory");
    }
}

__global__ void __launch_bounds__(kPairThreads, 1)
gemm_tf32x3_fused_astat_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                                    const __grid_constant__ CUtensorMap map_c, const Params p, int* raise_flag) {
    using namespace astat;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t panel_base = smem_base;
    const uint32_t ring_base = smem_base + kPanelBytes;
    const uint32_t store_base = ring_base + kBStages * kBStage;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_gen + kPanelBytes + kBStages * kBStage + kStoreBytes);
    const uint32_t bar_base = smem_u32(bars);
    auto a_full = [&](int kb) { return bar_base + 8u * kb; };
    auto a_ready = [&](int kb) { return bar_base + 8u * (kMaxKB + kb); };
    auto a_empty = [&](int kb) { return bar_base + 8u * (2 * kMaxKB + kb); };
    auto b_full = [&](int s) { return bar_base + 8u * (3 * kMaxKB + s); };
    auto b_empty = [&](int s) { return bar_base + 8u * (3 * kMaxKB + kBStages + s); };
    auto b_ready = [&](int s) { return bar_base + 8u * (3 * kMaxKB + 2 * kBStages + s); };
    auto tmem_full_bar = [&](int s) { return bar_base + 8u * (3 * kMaxKB + 3 * kBStages + s); };
    auto tmem_empty_bar = [&](int s) { return bar_base + 8u * (3 * kMaxKB + 3 * kBStages + 2 + s); };
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * kMaxKB + 3 * kBStages + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const uint32_t pair_id = blockIdx.x >> 1, num_pairs = gridDim.x >> 1;
    // Work unit = one A panel (b, mt) x one segment of its n-tiles; units are dealt round-robin in the order
    // (b, segment, mt), so the pairs that run side by side hold different panels of the SAME batch entry and
    // walk the same B tiles together (L2 hits).  p.nseg segments per panel (1 when there are panels enough).
    const uint32_t units = p.batch * p.tiles_m * p.nseg;
    auto unit_coords = [&](uint32_t u, uint32_t& b, uint32_t& mt, uint32_t& n_lo, uint32_t& n_hi) {
        mt = u % p.tiles_m;
        const uint32_t r = u / p.tiles_m;
        const uint32_t seg = r % p.nseg;
        b = r / p.nseg;
        n_lo = (uint32_t)(((uint64_t)seg * p.tiles_n) / p.nseg);
        n_hi = (uint32_t)(((uint64_t)(seg + 1) * p.tiles_n) / p.nseg);
    };
    // The i-th tile of a unit: every pair starts its panels at a DIFFERENT n-tile (rotation by pair_id * p.rot_mult; the
    // multiplier sends neighbouring pairs ~0.6 of the way round).  The pairs run at one pace, so without the rotation all of
    // them store the same 1 KiB column range of their 8 KiB rows of C at the same moment and load the same B tile; rotated,
    // the stores of any moment cover every column range.  Config 3, alternating processes on one box (scripts/exp/
    // exp_cfg3_ab.py, median ms): no rotation 1.260, rotation by pair_id 1.241, by 7 pair_id 1.240, by 3 pair_id 1.215, by
    // 5 pair_id 1.208-1.211 (-4 %); de-phasing the pairs in TIME on top (quarter-tile start delays) added 0.3 % and is not
    // done.  Tiles are independent: the order changes no bit of C.  TRN_GEMM_DEBUG bit 3 turns the rotation off.
    auto unit_tile = [&](uint32_t i, uint32_t n_lo, uint32_t n_hi) -> uint32_t {
        const uint32_t cnt = n_hi - n_lo;
        return n_lo + ((p.debug & 8u) ? i : (i + (pair_id * p.rot_mult) % cnt) % cnt);
    };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_a); tma_prefetch_desc(&map_b);
        for (int kb = 0; kb < kMaxKB; ++kb) {
            mbar_init(a_full(kb), 1);
            mbar_init(a_ready(kb), 2 * kSplitWarps);
            mbar_init(a_empty(kb), 1);
        }
        for (int s = 0; s < kBStages; ++s) {
            mbar_init(b_full(s), 1);
            mbar_init(b_empty(s), 1);
            mbar_init(b_ready(s), 2 * kSplitWarps);
        }
        for (int s = 0; s < 2; ++s) { mbar_init(tmem_full_bar(s), 1); mbar_init(tmem_empty_bar(s), 2 * kEpiWarps); }
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(tmem_slot)), "r"(kTmemCols) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    cluster_sync_all();
    if (p.trace && threadIdx.x == 0) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        p.trace[2 * blockIdx.x] = t;
    }
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 72;" ::: "memory");
    if (warp == 0) {
        // ===================== TMA producer (both CTAs) =====================
        if (elect_one()) {
            uint32_t stage = 0, phase = 0, panel = 0;
            for (uint32_t u = pair_id; u < units; u += num_pairs, ++panel) {
              uint32_t b, mt, n_lo, n_hi;
              unit_coords(u, b, mt, n_lo, n_hi);
              const uint32_t a_par = panel & 1u;
              for (uint32_t ti = 0; ti < n_hi - n_lo; ++ti) {
                const uint32_t nt = unit_tile(ti, n_lo, n_hi);
                const bool new_panel = ti == 0;
                const int m0 = (int)(mt * 2 * BM + rank * BM), n0 = (int)(nt * BN + rank * (BN / 2));
                for (uint32_t kb = 0; kb < p.num_kb; ++kb) {
                    const int k0 = (int)(kb * SBK);
                    if (new_panel) {
                        mbar_wait(a_empty(kb), a_par ^ 1);   // the previous panel's last tile is done with this k-block
                        mbar_expect_tx(a_full(kb), kABlk);
                        tma_load_3d(panel_base + kb * kASlot, &map_a, a_full(kb), k0, m0, (int)b);
                    }
                    mbar_wait(b_empty(stage), phase ^ 1);
                    const uint32_t sb = ring_base + stage * kBStage;
                    mbar_expect_tx(b_full(stage), kBRaw);
#pragma unroll
                    for (int j = 0; j < BN / 2 / 32; ++j)
                        tma_load_3d(sb + j * kBBox, &map_b, b_full(stage), n0 + 32 * j, k0, (int)b);
                    if (++stage == kBStages) { stage = 0; phase ^= 1; }
                }
              }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA only) =====================
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc_tf32(2 * BM, BN) | (1u << 16);   // B is MN-major
            uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0, panel = 0;
            for (uint32_t u = pair_id; u < units; u += num_pairs, ++panel) {
              uint32_t b, mt, n_lo, n_hi;
              unit_coords(u, b, mt, n_lo, n_hi);
              const uint32_t a_par = panel & 1u;
              for (uint32_t ti = 0; ti < n_hi - n_lo; ++ti) {
                const bool new_panel = ti == 0;
                const bool last_in_panel = ti + 1 == n_hi - n_lo;
                mbar_wait(tmem_empty_bar(acc), acc_phase ^ 1);   // K <= 128: one TMEM partial per tile
                tc_fence_after();
                const uint32_t d = tmem_base + acc * BN;
                for (uint32_t kb = 0; kb < p.num_kb; ++kb) {
                    if (new_panel) mbar_wait(a_ready(kb), a_par);
                    mbar_wait(b_ready(stage), phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t a_hi = panel_base + kb * kASlot, a_lo = a_hi + kABlk;
                        const uint32_t b_hi = ring_base + stage * kBStage, b_lo = b_hi + kBRaw;
#pragma unroll
                        for (int k = 0; k < SBK / UMMA_K; ++k) {
                            const uint32_t ka = k * UMMA_K * 4;
                            const uint32_t kbo = k * 1024;
                            const uint32_t accum = (kb | (uint32_t)k) != 0;
                            umma_tf32_pair(d, make_desc_k<SBK>(a_lo + ka), make_desc_mn_tf32(b_hi + kbo, kBBox), idesc, accum);
                            umma_tf32_pair(d, make_desc_k<SBK>(a_hi + ka), make_desc_mn_tf32(b_lo + kbo, kBBox), idesc, 1u);
                            umma_tf32_pair(d, make_desc_k<SBK>(a_hi + ka), make_desc_mn_tf32(b_hi + kbo, kBBox), idesc, 1u);
                        }
                        umma_commit_pair(b_empty(stage));
                        if (last_in_panel) umma_commit_pair(a_empty(kb));
                        if (kb + 1 == p.num_kb) umma_commit_pair(tmem_full_bar(acc));
                    }
                    __syncwarp();
                    if (++stage == kBStages) { stage = 0; phase ^= 1; }
                }
                if (++acc == 2) { acc = 0; acc_phase ^= 1; }
              }
            }
        }
    } else {
        // ===================== splitter warps (2-3, both CTAs) =====================
        const uint32_t sw = (uint32_t)warp - 2;
        bool bad = false;
        uint32_t stage = 0, phase = 0, panel = 0;
        for (uint32_t u = pair_id; u < units; u += num_pairs, ++panel) {
          uint32_t b, mt, n_lo, n_hi;
          unit_coords(u, b, mt, n_lo, n_hi);
          const uint32_t a_par = panel & 1u;
          for (uint32_t ti = 0; ti < n_hi - n_lo; ++ti) {
            const bool new_panel = ti == 0;
            for (uint32_t kb = 0; kb < p.num_kb; ++kb) {
                if (new_panel) {
                    mbar_wait(a_full(kb), a_par);
                    const uint32_t sa = panel_base + kb * kASlot;
                    split_block_8k(sa, sa + kABlk, sw, lane, bad);
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) mbar_arrive_cluster(a_ready(kb) & kPeerMask);
                }
                mbar_wait(b_full(stage), phase);
                const uint32_t sb = ring_base + stage * kBStage;
                split_block_8k(sb, sb + kBRaw, sw, lane, bad);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(b_ready(stage) & kPeerMask);
                if (++stage == kBStages) { stage = 0; phase ^= 1; }
            }
          }
        }
        if (bad) *raise_flag = 1;
    }
    } else {
        // ===================== epilogue (warps 4..11, both CTAs, own 128 rows) =====================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;" ::: "memory");
        uint32_t acc = 0, acc_phase = 0;
        for (uint32_t u = pair_id; u < units; u += num_pairs) {
            uint32_t b, mt, n_lo, n_hi;
            unit_coords(u, b, mt, n_lo, n_hi);
            for (uint32_t ti = 0; ti < n_hi - n_lo; ++ti)
                pair_epilogue_tile(&map_c, p, tmem_base, store_base, tmem_full_bar(0), tmem_empty_bar(0), 1u, acc, acc_phase,
                                   b, mt, unit_tile(ti, n_lo, n_hi), warp - 4, lane, rank);
        }
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
        if (p.trace && warp == 4 && lane == 0) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            p.trace[2 * blockIdx.x + 1] = t;
        }
    }

    tc_fence_before();
    cluster_sync_all();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" :: "r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---- operand pre-pass: split into tf32 hi/lo, pad K, lay B out K-major ------------------------------
// returns true when x is Inf/NaN (the caller raises the fallback flag)
__device__ __forceinline__ bool split_tf32(float x, float& hi, float& lo) {
    if (!isfinite(x)) { hi = x; lo = 0.f; return true; }
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    float fh = __uint_as_float(h);
    if (isinf(fh)) fh = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);  // rounding overflowed: truncate instead
    const float r = x - fh;                              // exact
    uint32_t l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
    hi = fh;
    lo = __uint_as_float(l);
    return false;
}

// The split kernels run on flat grids (one tile per CTA): measured on B200, flat streaming grids reach the
// HBM roofline where persistent grid-stride loops of the same body stay 10-25 % below it (scripts/exp/exp_map.cu).

// rows x k (row-major, per batch) -> hi/lo [rows x kpad]; 2048 elements per CTA
__global__ void __launch_bounds__(256)
split_rows_kernel(const float* __restrict__ in, float* __restrict__ hi, float* __restrict__ lo,
                  size_t rows_total, size_t k, size_t kpad, int* __restrict__ flag) {
    const size_t total = rows_total * kpad;
    bool bad = false;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
        const size_t i = (size_t)blockIdx.x * 2048 + u * 256 + threadIdx.x;
        if (i < total) {
            const size_t r = i / kpad, c = i - r * kpad;
            float h = 0.f, l = 0.f;
            if (c < k) bad |= split_tf32(ld_stream(in + r * k + c), h, l);
            hi[i] = h;
            lo[i] = l;
        }
    }
    if (bad) *flag = 1;
}
// 4-wide variant for k % 4 == 0 (kpad is always a multiple of 32); 512 float4 per CTA
__global__ void __launch_bounds__(256)
split_rows_vec_kernel(const float* __restrict__ in, float* __restrict__ hi, float* __restrict__ lo,
                      size_t rows_total, size_t k, size_t kpad, int* __restrict__ flag) {
    const size_t kv = k >> 2, kpv = kpad >> 2;
    const size_t total = rows_total * kpv;
    bool bad = false;
    float4 x[2];
    size_t idx[2];
    bool live[2];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        idx[u] = (size_t)blockIdx.x * 512 + u * 256 + threadIdx.x;
        const size_t r = idx[u] / kpv, c = idx[u] - r * kpv;
        live[u] = idx[u] < total && c < kv;
        x[u] = live[u] ? ld_stream(reinterpret_cast<const float4*>(in + r * k) + c) : make_float4(0, 0, 0, 0);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        if (idx[u] < total) {
            float4 h = make_float4(0, 0, 0, 0), l = h;
            if (live[u]) {
                bad |= split_tf32(x[u].x, h.x, l.x); bad |= split_tf32(x[u].y, h.y, l.y);
                bad |= split_tf32(x[u].z, h.z, l.z); bad |= split_tf32(x[u].w, h.w, l.w);
            }
            st_stream(reinterpret_cast<float4*>(hi) + idx[u], h);
            st_stream(reinterpret_cast<float4*>(lo) + idx[u], l);
        }
    }
    if (bad) *flag = 1;
}
// B [batch][k][n] row-major -> hi/lo [batch][n][kpad]: one 64(k) x 64(n) tile per CTA through shared memory,
// 128-bit accesses on the write side always and on the read side when rows of B are 16-byte aligned (VEC).
template <bool VEC>
__global__ void __launch_bounds__(256)
split_transpose_kernel(const float* __restrict__ in, float* __restrict__ hi, float* __restrict__ lo,
                       size_t batch, size_t k, size_t n, size_t kpad, int* __restrict__ flag) {
    __shared__ float tile[64][65];   // [k][n]
    const size_t tiles_n = (n + 63) / 64, tiles_k = (kpad + 63) / 64;
    const size_t per_batch = tiles_n * tiles_k;
    const size_t t = blockIdx.x;
    const size_t b = t / per_batch, r = t % per_batch;
    const size_t tk = r / tiles_n, tn = r % tiles_n;
    const float* src = in + b * k * n;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int idx = it * 256 + threadIdx.x;
        const int kr = idx >> 4, c4 = (idx & 15) * 4;
        const size_t kk = tk * 64 + kr, nn = tn * 64 + c4;
        float4 v = make_float4(0, 0, 0, 0);
        if (kk < k) {
            if (VEC) {
                if (nn < n) v = ld_stream(reinterpret_cast<const float4*>(src + kk * n + nn));
            } else {
                if (nn < n) v.x = ld_stream(src + kk * n + nn);
                if (nn + 1 < n) v.y = ld_stream(src + kk * n + nn + 1);
                if (nn + 2 < n) v.z = ld_stream(src + kk * n + nn + 2);
                if (nn + 3 < n) v.w = ld_stream(src + kk * n + nn + 3);
            }
        }
        tile[kr][c4] = v.x; tile[kr][c4 + 1] = v.y; tile[kr][c4 + 2] = v.z; tile[kr][c4 + 3] = v.w;
    }
    __syncthreads();
    bool bad = false;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        const int idx = it * 256 + threadIdx.x;
        const int nr = idx >> 4, k4 = (idx & 15) * 4;
        const size_t nn = tn * 64 + nr, kk = tk * 64 + k4;
        if (nn < n && kk < kpad) {
            float4 h, l;
            bad |= split_tf32(tile[k4][nr], h.x, l.x);
            bad |= split_tf32(tile[k4 + 1][nr], h.y, l.y);
            bad |= split_tf32(tile[k4 + 2][nr], h.z, l.z);
            bad |= split_tf32(tile[k4 + 3][nr], h.w, l.w);
            const size_t o = (b * n + nn) * kpad + kk;
            st_stream(reinterpret_cast<float4*>(hi + o), h);
            st_stream(reinterpret_cast<float4*>(lo + o), l);
        }
    }
    if (bad) *flag = 1;
}

// lo half only, for an operand whose RAW words serve as the hi half: kind::tf32 reads the top 19 bits of a word, i.e. the
// tensor core itself uses hi = x & 0xFFFFE000 (as in the fused-split kernels), and lo = tf32_rna(x - hi) is all a pre-pass has
// to write — one read and ONE write per element instead of two.  Needs k % 32 == 0 (no K padding) and 16-byte aligned data.
__global__ void __launch_bounds__(256)
split_rows_lo_kernel(const float* __restrict__ in, float* __restrict__ lo, size_t nvec, int* __restrict__ flag) {
    const float4* in4 = reinterpret_cast<const float4*>(in);
    float4* lo4 = reinterpret_cast<float4*>(lo);
    bool bad = false;
    auto one = [&](float x) -> float {
        if (!isfinite(x)) { bad = true; return 0.f; }
        const float r = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);   // exact
        uint32_t l;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(r));
        return __uint_as_float(l);
    };
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const size_t i = (size_t)blockIdx.x * 512 + u * 256 + threadIdx.x;
        if (i < nvec) {
            const float4 x = ld_stream(in4 + i);
            lo4[i] = make_float4(one(x.x), one(x.y), one(x.z), one(x.w));
        }
    }
    if (bad) *flag = 1;
}

// ---- host side ---------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (EncodeTiledFn)p;
    return fn;
}

// [batch][rows][kpad] f32, box = {box_cols, box_rows, 1}; swizzle span = box_cols * 4 bytes (128 or 64)
int make_map(CUtensorMap* map, const float* base, size_t batch, size_t rows, size_t kpad, uint32_t box_rows,
             uint32_t box_cols, bool atom32) {
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail(TRN_GPU_ERROR, "cuTensorMapEncodeTiled is not available from the CUDA driver");
    cuuint64_t dims[3] = {kpad, rows, batch};
    cuuint64_t strides[2] = {kpad * sizeof(float), rows * kpad * sizeof(float)};
    cuuint32_t box[3] = {box_cols, box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE,
                     atom32 ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : box_cols == 16 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(TRN_GPU_ERROR, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return TRN_OK;
}

template <int TERMS, int SBK>
static int launch(const CUtensorMap& ah, const CUtensorMap& al, const CUtensorMap& bh, const CUtensorMap& bl,
                  const CUtensorMap& mc, const Params& p, int sm_count, cudaStream_t s) {
    using C = Cfg<TERMS, SBK>;
    // thread-safe one-time opt-in to > 48 KiB of dynamic shared memory
    static const cudaError_t attr = cudaFuncSetAttribute(gemm_tf32_kernel<TERMS, SBK>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::kSmemBytes);
    TRN_CUDA(attr);
    const uint32_t total = p.tiles_m * p.tiles_n * p.batch;
    const uint32_t grid = total < (uint32_t)sm_count ? total : (uint32_t)sm_count;
    gemm_tf32_kernel<TERMS, SBK><<<grid, kThreads, C::kSmemBytes, s>>>(ah, al, bh, bl, mc, p);
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

}  // namespace tc

// ---- optional live timing of the pre-pass and the main kernel (trn_profile_*) -------------------------
// A ring of event triples so a whole timed region of back-to-back calls can be profiled without a host sync
// between them; trn_profile_last_gemm() averages over the calls recorded since trn_profile_enable(1).
constexpr int kProfRing = 64;
static bool g_profile = false;
static cudaEvent_t g_ev[kProfRing][3];
static bool g_ev_created = false;
static unsigned g_ev_count = 0;

// brackets of one profiled GEMM call: begin (before the pre-pass), mid (pre-pass done), end (main kernel done)
static cudaEvent_t* profile_slot() {
    if (!g_profile) return nullptr;
    if (!g_ev_created) {
        for (auto& t : g_ev) for (auto& e : t) if (cudaEventCreate(&e) != cudaSuccess) return nullptr;
        g_ev_created = true;
    }
    return g_ev[g_ev_count % kProfRing];
}
void gemm_profile_begin(cudaStream_t s) { if (cudaEvent_t* ev = profile_slot()) cudaEventRecord(ev[0], s); }
void gemm_profile_mid(cudaStream_t s) { if (cudaEvent_t* ev = profile_slot()) cudaEventRecord(ev[1], s); }
void gemm_profile_end(cudaStream_t s) {
    if (cudaEvent_t* ev = profile_slot()) {
        cudaEventRecord(ev[2], s);
        ++g_ev_count;
    }
}

bool gemm_tc_supported(size_t m, size_t k, size_t n) {
    // 32-bit tile coordinates and TMA dimension limits; any m, n, k >= 1 otherwise
    return m >= 1 && n >= 1 && k >= 1 && m < (1u << 30) && n < (1u << 30) && k < (1u << 30);
}

size_t gemm_tc_kpad(size_t k) { return (k + tc::BK - 1) / tc::BK * tc::BK; }

// A [batch][m][k] row-major -> hi/lo [batch][m][kpad]; raises *flag on Inf/NaN
int gemm_tc_split_a(const float* a, float* a_hi, float* a_lo, size_t batch, size_t m, size_t k, int* flag,
                    cudaStream_t s) {
    using namespace tc;
    Context* cx = ctx();
    if (!cx) return TRN_GPU_ERROR;
    const size_t kpad = gemm_tc_kpad(k);
    const size_t a_elems = batch * m * kpad;
    if (a_elems == 0) return TRN_OK;
    const bool vec = (k % 4 == 0) && ((reinterpret_cast<uintptr_t>(a) & 15u) == 0);
    const size_t blocks = vec ? (a_elems / 4 + 511) / 512 : (a_elems + 2047) / 2048;
    if (blocks > 0x7FFFFFFFull) return fail(TRN_INVALID_INPUT, "operand of %zu elements exceeds the launch grid", a_elems);
    if (vec) split_rows_vec_kernel<<<(unsigned)blocks, 256, 0, s>>>(a, a_hi, a_lo, batch * m, k, kpad, flag);
    else     split_rows_kernel<<<(unsigned)blocks, 256, 0, s>>>(a, a_hi, a_lo, batch * m, k, kpad, flag);
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

bool gemm_tc_raw_hi_ok(const float* a, size_t k) { return k % tc::BK == 0 && (reinterpret_cast<uintptr_t>(a) & 15u) == 0; }
int gemm_tc_split_a_lo(const float* a, float* a_lo, size_t batch, size_t m, size_t k, int* flag, cudaStream_t s) {
    if (!ctx()) return TRN_GPU_ERROR;
    const size_t nvec = batch * m * k / 4;
    if (nvec == 0) return TRN_OK;
    const size_t blocks = (nvec + 511) / 512;
    if (blocks > 0x7FFFFFFFull) return fail(TRN_INVALID_INPUT, "operand of %zu elements exceeds the launch grid", nvec * 4);
    tc::split_rows_lo_kernel<<<(unsigned)blocks, 256, 0, s>>>(a, a_lo, nvec, flag);
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

// B [batch][k][n] row-major -> hi/lo [batch][n][kpad] (K-major); raises *flag on Inf/NaN
int gemm_tc_split_b(const float* b, float* b_hi, float* b_lo, size_t batch, size_t k, size_t n, int* flag,
                    cudaStream_t s) {
    using namespace tc;
    Context* cx = ctx();
    if (!cx) return TRN_GPU_ERROR;
    const size_t kpad = gemm_tc_kpad(k);
    if (batch * n * kpad == 0) return TRN_OK;
    const size_t tiles = batch * ((n + 63) / 64) * ((kpad + 63) / 64);
    if (tiles > 0x7FFFFFFFull) return fail(TRN_INVALID_INPUT, "operand of %zu tiles exceeds the launch grid", tiles);
    const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(b) & 15u) == 0);
    if (vec) split_transpose_kernel<true><<<(unsigned)tiles, 256, 0, s>>>(b, b_hi, b_lo, batch, k, n, kpad, flag);
    else     split_transpose_kernel<false><<<(unsigned)tiles, 256, 0, s>>>(b, b_hi, b_lo, batch, k, n, kpad, flag);
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

// C[b] = A[b] * B[b] from pre-split operands.  The tensor-core kernel returns at once when *flag != 0.
int gemm_tc_main(const float* a_hi, const float* a_lo, const float* b_hi, const float* b_lo, float* c, size_t batch,
                 size_t m, size_t k, size_t n, int terms, const int* flag, cudaStream_t s, size_t route_m) {
    using namespace tc;
    Context* cx = ctx();
    if (!cx) return TRN_GPU_ERROR;
    if (batch == 0 || m == 0 || n == 0) return TRN_OK;
    const size_t kpad = gemm_tc_kpad(k);
    // CTA-pair kernel (cta_group::2) for 3xTF32 whenever there are at least two 128-row tiles of C;
    // TRN_GEMM_PAIR=0 forces the single-CTA kernel (A/B measurements, and its own parity tests).
    static const int use_pair = [] { const char* e = getenv("TRN_GEMM_PAIR"); return e ? atoi(e) : 1; }();
    if (terms == 3 && use_pair && (route_m ? route_m : m) > (size_t)BM) {   // (a short row block of a tall product keeps the pair kernel)
        CUtensorMap pa_h, pa_l, pb_h, pb_l, pc;
        TRN_TRY(make_map(&pa_h, a_hi, batch, m, kpad, BM, pair::SBK));
        TRN_TRY(make_map(&pa_l, a_lo, batch, m, kpad, BM, pair::SBK));
        TRN_TRY(make_map(&pb_h, b_hi, batch, n, kpad, BN / 2, pair::SBK));
        TRN_TRY(make_map(&pb_l, b_lo, batch, n, kpad, BN / 2, pair::SBK));
        pc = pa_h;
        const bool tma_store = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(c) & 15u) == 0);
        if (tma_store) TRN_TRY(make_map(&pc, c, batch, m, n, 32, 32));
        Params p;
        p.c = c;
        p.tma_store = tma_store ? 1u : 0u;
        p.nonfinite_flag = flag;
        p.m = (uint32_t)m;
        p.n = (uint32_t)n;
        p.num_kb = (uint32_t)(kpad / pair::SBK);
        p.tiles_m = (uint32_t)((m + 2 * BM - 1) / (2 * BM));
        p.tiles_n = (uint32_t)((n + BN - 1) / BN);
        p.batch = (uint32_t)batch;
        static const cudaError_t smem_optin = cudaFuncSetAttribute(gemm_tf32x3_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, pair::kSmemBytes);
        TRN_CUDA(smem_optin);
        const uint32_t total = p.tiles_m * p.tiles_n * p.batch;
        const uint32_t max_pairs = (uint32_t)cx->sm_count / 2;
        const uint32_t pairs = total < max_pairs ? total : max_pairs;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * pairs);
        cfg.blockDim = dim3(kPairThreads);
        cfg.dynamicSmemBytes = pair::kSmemBytes;
        cfg.stream = s;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        TRN_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32x3_pair_kernel, pa_h, pa_l, pb_h, pb_l, pc, p));
        count_launch();
        return TRN_OK;
    }
    // K extent of one pipeline stage (see Cfg): 16 for 3xTF32 (4-deep ring), 32 for the 1xTF32 probe.
    // TRN_GEMM_STAGE_K=32 restores the 2 x 96 KiB ring for A/B measurements.
    static const int stage_k_3x = [] { const char* e = getenv("TRN_GEMM_STAGE_K"); return (e && atoi(e) == 32) ? 32 : 16; }();
    const uint32_t sbk = terms == 3 ? (uint32_t)stage_k_3x : 32u;
    CUtensorMap map_ah, map_al, map_bh, map_bl;
    TRN_TRY(make_map(&map_ah, a_hi, batch, m, kpad, BM, sbk));
    TRN_TRY(make_map(&map_al, a_lo, batch, m, kpad, BM, sbk));
    TRN_TRY(make_map(&map_bh, b_hi, batch, n, kpad, BN, sbk));
    TRN_TRY(make_map(&map_bl, b_lo, batch, n, kpad, BN, sbk));

    // C as a [batch][m][n] tensor with 32x32 store boxes — only when its rows are 16-byte aligned
    CUtensorMap map_c = map_ah;
    const bool tma_store = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(c) & 15u) == 0);
    if (tma_store) TRN_TRY(make_map(&map_c, c, batch, m, n, 32, 32));

    Params p;
    p.c = c;
    p.tma_store = tma_store ? 1u : 0u;
    p.nonfinite_flag = flag;
    p.m = (uint32_t)m;
    p.n = (uint32_t)n;
    p.num_kb = (uint32_t)(kpad / sbk);
    p.tiles_m = (uint32_t)((m + BM - 1) / BM);
    p.tiles_n = (uint32_t)((n + BN - 1) / BN);
    p.batch = (uint32_t)batch;
    if (terms == 3)
        return sbk == 16 ? launch<3, 16>(map_ah, map_al, map_bh, map_bl, map_c, p, cx->sm_count, s)
                         : launch<3, 32>(map_ah, map_al, map_bh, map_bl, map_c, p, cx->sm_count, s);
    return launch<1, 32>(map_ah, map_al, map_bh, map_bl, map_c, p, cx->sm_count, s);
}

// The fused-split kernel's preconditions (TMA strides of the RAW operands) and the shapes it is the default for.
// TRN_GEMM_FUSED = 0 / 1 forces the pre-pass / the fused kernel for A/B measurements and the parity tests.
bool gemm_tc_fused_ok(const float* a, const float* b, size_t m, size_t k, size_t n) {
    return m > (size_t)tc::BM && k % 4 == 0 && n % 4 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15u) == 0;
}
static int fused_mode() {
    static const int mode = [] { const char* e = getenv("TRN_GEMM_FUSED"); return e ? atoi(e) : -1; }();
    return mode;
}
static size_t fused_max_k() {
    static const size_t v = [] { const char* e = getenv("TRN_GEMM_FUSED_MAXK"); return e ? (size_t)atol(e) : (size_t)384; }();
    return v;
}
bool gemm_tc_uses_fused(const float* a, const float* b, size_t m, size_t k, size_t n) {
    if (!gemm_tc_fused_ok(a, b, m, k, n)) return false;
    const int mode = fused_mode();
    return mode == 1 || (mode != 0 && k <= fused_max_k());
}

// C[b] = A[b] * B[b] straight from the raw operands (no scratch): the splitter warps raise *flag on Inf/NaN.
int gemm_tc_fused_main(const float* a, const float* b, float* c, size_t batch, size_t m, size_t k, size_t n, int* flag,
                       cudaStream_t s) {
    using namespace tc;
    Context* cx = ctx();
    if (!cx) return TRN_GPU_ERROR;
    if (batch == 0 || m == 0 || n == 0) return TRN_OK;
    CUtensorMap ma, mb, mc;
    TRN_TRY(make_map(&ma, a, batch, m, k, BM, fused::SBK));     // [batch][m][k], boxes of 128 rows x 16 k, SWIZZLE_64B
    TRN_TRY(make_map(&mb, b, batch, k, n, fused::SBK, 32, true));   // [batch][k][n], boxes of 16 k x 32 n, SWIZZLE_128B_ATOM_32B
    mc = ma;
    const bool tma_store = (reinterpret_cast<uintptr_t>(c) & 15u) == 0;   // n % 4 == 0 is a precondition here
    if (tma_store) TRN_TRY(make_map(&mc, c, batch, m, n, 32, 32));
    Params p;
    p.c = c;
    p.tma_store = tma_store ? 1u : 0u;
    p.nonfinite_flag = flag;
    p.m = (uint32_t)m;
    p.n = (uint32_t)n;
    p.num_kb = (uint32_t)((k + fused::SBK - 1) / fused::SBK);
    p.tiles_m = (uint32_t)((m + 2 * BM - 1) / (2 * BM));
    p.tiles_n = (uint32_t)((n + BN - 1) / BN);
    p.batch = (uint32_t)batch;
    static const uint32_t terms_mask = [] { const char* e = getenv("TRN_GEMM_FUSED_TERMS"); const int v = e ? atoi(e) : 7; return (uint32_t)(v >= 1 && v <= 7 ? v : 7); }();
    p.terms_mask = terms_mask;
    static const uint32_t debug_bits = [] { const char* e = getenv("TRN_GEMM_DEBUG"); return (uint32_t)(e ? atoi(e) : 0); }();
    p.debug = debug_bits;
    static const int hint_on = [] { const char* e = getenv("TRN_GEMM_STORE_HINT"); return e ? atoi(e) : 1; }();
    p.store_hint = (hint_on && batch * m * n * sizeof(float) > ((size_t)64 << 20)) ? 1u : 0u;
    static const int trace_on = [] { const char* e = getenv("TRN_GEMM_TRACE"); return e ? atoi(e) : 0; }();
    static unsigned long long* trace_buf = nullptr;
    if (trace_on && !trace_buf) cudaMalloc(&trace_buf, 2 * 1024 * sizeof(unsigned long long));
    p.trace = trace_on ? trace_buf : nullptr;
    static const cudaError_t smem_optin = cudaFuncSetAttribute(gemm_tf32x3_fused_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, fused::kSmemBytes);
    TRN_CUDA(smem_optin);
    static const cudaError_t smem_optin2 = cudaFuncSetAttribute(gemm_tf32x3_fused_astat_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, astat::kSmemBytes);
    TRN_CUDA(smem_optin2);
    // K <= 128 and at least two n-tiles per row of C: keep the A panel in shared memory (TRN_GEMM_ASTAT=0 disables)
    static const int astat_on = [] { const char* e = getenv("TRN_GEMM_ASTAT"); return e ? atoi(e) : 1; }();
    const bool use_astat = astat_on && p.num_kb <= (uint32_t)astat::kMaxKB && p.tiles_n >= 2;
    const uint32_t max_pairs_hw = (uint32_t)cx->sm_count / 2;
    {   // segments per A panel: the split that minimises the busiest pair's work — rounds of units x (tiles per unit + a
        // panel-switch allowance); config 3 (2048 panels) keeps whole panels, one 8192 x 128 x 8192 product gets 16 segments
        const uint64_t panels = (uint64_t)p.batch * p.tiles_m;
        double best = 1e300;
        uint32_t best_nseg = 1;
        for (uint32_t nseg = 1; nseg <= p.tiles_n; ++nseg) {
            const uint64_t units = panels * nseg;
            const uint64_t rounds = (units + max_pairs_hw - 1) / max_pairs_hw;
            const uint32_t tpu = (p.tiles_n + nseg - 1) / nseg;
            const double cost = (double)rounds * ((double)tpu + 0.35);
            if (cost < best - 1e-9) { best = cost; best_nseg = nseg; }
        }
        p.nseg = best_nseg;
        // rotation multiplier: the integer nearest 0.618 x (tiles per unit) that is coprime with it (8 tiles -> 5)
        const uint32_t tpu = (p.tiles_n + p.nseg - 1) / p.nseg;
        auto gcd = [](uint32_t a, uint32_t b) { while (b) { const uint32_t t = a % b; a = b; b = t; } return a; };
        uint32_t mult = 1;
        if (tpu > 2) {
            const uint32_t want = (uint32_t)(0.618 * tpu + 0.5);
            for (uint32_t d = 0; d < tpu; ++d) {
                if (want + d < tpu && gcd(want + d, tpu) == 1) { mult = want + d; break; }
                if (want > d && gcd(want - d, tpu) == 1) { mult = want - d; break; }
            }
        }
        p.rot_mult = mult;
    }
    const uint64_t total = use_astat ? (uint64_t)p.tiles_m * p.batch * p.nseg : (uint64_t)p.tiles_m * p.tiles_n * p.batch;
    const uint32_t max_pairs = (uint32_t)cx->sm_count / 2;
    const uint32_t pairs = total < max_pairs ? (uint32_t)total : max_pairs;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * pairs);
    cfg.blockDim = dim3(kPairThreads);
    cfg.dynamicSmemBytes = use_astat ? astat::kSmemBytes : fused::kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (use_astat) TRN_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32x3_fused_astat_pair_kernel, ma, mb, mc, p, flag));
    else TRN_CUDA(cudaLaunchKernelEx(&cfg, gemm_tf32x3_fused_pair_kernel, ma, mb, mc, p, flag));
    count_launch();
    if (p.trace && use_astat && trace_on > 1) {   // experiments only (TRN_GEMM_TRACE=2): when did every CTA finish?
        cudaStreamSynchronize(s);
        std::vector<unsigned long long> h(4 * pairs);
        cudaMemcpy(h.data(), trace_buf, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
        unsigned long long t0 = ~0ull;
        for (uint32_t i = 0; i < 2 * pairs; ++i) t0 = h[2 * i] < t0 ? h[2 * i] : t0;
        std::vector<double> st, en;
        for (uint32_t i = 0; i < 2 * pairs; ++i) { st.push_back((h[2 * i] - t0) * 1e-3); en.push_back((h[2 * i + 1] - t0) * 1e-3); }
        std::sort(st.begin(), st.end()); std::sort(en.begin(), en.end());
        fprintf(stderr, "[gemm trace] %u CTAs: start max %.1f us; end min %.1f p10 %.1f median %.1f p90 %.1f max %.1f us\n", 2 * pairs,
                st.back(), en.front(), en[en.size() / 10], en[en.size() / 2], en[en.size() * 9 / 10], en.back());
    }
    return TRN_OK;
}

int launch_gemm_tc(const float* a, const float* b, float* c, size_t batch, size_t m, size_t k, size_t n, int terms,
                   cudaStream_t s, size_t route_m) {
    Context* cx = ctx();
    if (!cx) return TRN_GPU_ERROR;
    if (batch == 0 || m == 0 || n == 0) return TRN_OK;
    const size_t kpad = gemm_tc_kpad(k);
    const size_t a_elems = batch * m * kpad, b_elems = batch * n * kpad;

    if (terms == 3 && gemm_tc_uses_fused(a, b, route_m ? route_m : m, k, n)) {
        // no pre-pass, no operand scratch: only the 4-byte non-finite flag the splitter warps may raise
        int* flag = nullptr;
        TRN_TRY(scratch_alloc((void**)&flag, 256, s));
        TRN_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));
        gemm_profile_begin(s);
        gemm_profile_mid(s);
        int st = gemm_tc_fused_main(a, b, c, batch, m, k, n, flag, s);
        gemm_profile_end(s);
        if (st == TRN_OK) st = launch_gemm_simt(a, b, c, batch, m, k, n, s, flag);
        scratch_free(flag, s);
        return st;
    }

    // scratch: A_hi, A_lo, Bt_hi, Bt_lo (stream-ordered pool; stays cached between calls)
    float* scratch = nullptr;
    TRN_TRY(scratch_alloc((void**)&scratch, (2 * a_elems + 2 * b_elems) * sizeof(float) + 256, s));
    float* a_hi = scratch;
    float* a_lo = a_hi + a_elems;
    float* b_hi = a_lo + a_elems;
    float* b_lo = b_hi + b_elems;
    int* flag = reinterpret_cast<int*>(b_lo + b_elems);   // non-finite-input flag (see header)
    TRN_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));

    gemm_profile_begin(s);
    int st = gemm_tc_split_a(a, a_hi, a_lo, batch, m, k, flag, s);
    if (st == TRN_OK) st = gemm_tc_split_b(b, b_hi, b_lo, batch, k, n, flag, s);
    gemm_profile_mid(s);
    if (st == TRN_OK) st = gemm_tc_main(a_hi, a_lo, b_hi, b_lo, c, batch, m, k, n, terms, flag, s, route_m);
    gemm_profile_end(s);
    // IEEE fallback for Inf/NaN inputs: runs only when the flag is set (checked on the device)
    if (st == TRN_OK) st = launch_gemm_simt(a, b, c, batch, m, k, n, s, flag);
    scratch_free(scratch, s);
    return st;
}

}  // namespace trn

extern "C" {
int trn_profile_enable(int on) {
    trn::g_profile = on != 0;
    if (on) trn::g_ev_count = 0;
    return TRN_OK;
}
int trn_profile_last_gemm(float* prepass_ms, float* kernel_ms) {
    using namespace trn;
    if (g_ev_count == 0) return fail(TRN_INVALID_INPUT, "no profiled GEMM call has been made");
    const unsigned cnt = g_ev_count < (unsigned)kProfRing ? g_ev_count : (unsigned)kProfRing;
    TRN_CUDA(cudaEventSynchronize(g_ev[(g_ev_count - 1) % kProfRing][2]));
    float a = 0.f, b = 0.f;
    for (unsigned i = 0; i < cnt; ++i) {
        float x = 0.f, y = 0.f;
        TRN_CUDA(cudaEventElapsedTime(&x, g_ev[i][0], g_ev[i][1]));
        TRN_CUDA(cudaEventElapsedTime(&y, g_ev[i][1], g_ev[i][2]));
        a += x;
        b += y;
    }
    a /= (float)cnt;
    b /= (float)cnt;
    if (prepass_ms) *prepass_ms = a;
    if (kernel_ms) *kernel_ms = b;
    return TRN_OK;
}
}
