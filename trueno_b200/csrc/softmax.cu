// softmax.cu — row softmax / log_softmax with ONE read and ONE write of HBM per element.
//
// Replaces Vector::softmax / log_softmax (src/vector.rs:1516-1553, :1581-1623: max -> expf(x-max)
// -> sum -> divide, i.e. 3 reads + 2 writes per element on the CPU, 4 dispatches with a host round
// trip each in the wgpu path, src/backends/gpu/device.rs:956-971).
//
// Kernels, chosen by row length; measured on B200 in scripts/sweep_rows.py, scripts/sweep_rows_ext.py:
//   * cols <= 1024: one WARP per row, 8 independent 128-bit loads in flight per lane, shuffle-only statistics,
//     flat grid — 6.8 TB/s from 128 to 1024 columns (1.03 of the measured copy bandwidth).
//   * 1024 < cols <= 16384: the row in the REGISTERS of one CTA (256 or 512 threads x up to 8 float4), one row
//     per CTA on a flat grid — 6.7-6.9 TB/s.
//   * aligned rows of 28672 < cols <= 32768 (config 5: 32 000): the TMA RING kernel.  One persistent CTA per SM;
//     a producer warp streams rows into a 224 KiB shared-memory ring (7 x 32 KiB or 14 x 16 KiB slots) with 1-D bulk
//     copies (cp.async.bulk + mbarrier complete_tx), 16 consumer warps pull each slot into registers
//     (a row = 16 float4 per thread), release the slot at once, and compute max -> exp -> sum -> scale out of
//     registers.  Because slots are released as soon as they are in registers, the next row's bulk copies run
//     underneath the current row's exp and store phases — HBM reads never stop, with no register cost.  6.0 TB/s (0.91).
//   * every other row longer than 16384 columns (LLM-vocabulary rows: 32 001, 50 257, 128 256, 151 936, 262 144 ...):
//     the TWO-PASS cluster kernel — a row spread over a 1- / 2- / 4- / 8-CTA cluster: pass 1 streams the row from HBM
//     keeping an online (max, sum) pair per thread (one rescale per 16 elements) and parks it in L2 (evict_last hint),
//     the pairs are folded through the block and through distributed shared memory in a fixed order; pass 2 re-reads
//     the row from L2 (evict_first) and writes the result.  HBM sees one read and one write per element.  5.2-6.2 TB/s.
//   * a FEW long rows (one large Vector::softmax): every row split over many CTAs, two launches (segment pairs to
//     a workspace, then fold + write) — the whole machine works on one vector.
//   * rows that are not 16-byte aligned (cols % 4 != 0: 77, 1001, 50 257 ...; or a base pointer that is only
//     4-byte aligned): the same kernels in their WINDOW form — the 16-byte aligned body of a row streams exactly
//     as above and the up to 3 + 3 elements before / after it ride in six designated threads.
// Only input / output pointers whose misalignment differs take the three-pass fallback kernel (TRN_ROWS_GENERIC=1
// forces it, for A/B tests).  All reductions use fixed trees, so reruns are bit-identical.
// Math: accurate expf / logf; softmax scales by the correctly rounded reciprocal of the row sum
// (<= 1 ulp from the reference's e / sum) — no fast-math intrinsics.
//
// Algorithmic bytes per element: 8 B (4 read + 4 written).  HBM-bound.
#include <cooperative_groups.h>
#include <stdlib.h>

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


// A row as a 16-byte aligned BODY of 128-bit vectors plus (WIN only) up to 3 + 3 edge elements.
// WIN = false: the row starts on a 16-byte boundary and cols % 4 == 0 — the body is the row.  WIN = true: the row
// starts `mis` elements (0..3) past a 16-byte boundary and / or cols % 4 != 0: the body starts at the first aligned
// element, `head` = (4 - mis) & 3 elements precede it and `tail` = (cols - head) & 3 follow it.  Every kernel streams
// the body exactly as in the aligned case and lets six designated threads carry one edge element each through the
// same max / sum / normalise steps, so nothing outside [row, row + cols) is read or written and the body loop has no
// per-vector tests.  Input and output share `mis` (checked by the dispatcher).
template <bool WIN>
struct RowView {
    const float4* vsrc;
    float4* vdst;
    size_t nvec;           // body vectors
    const float* src;
    float* dst;
    size_t cols;
    unsigned head, tail;
    __device__ __forceinline__ RowView(const float* in, float* out, size_t row, size_t cols_, unsigned mis0)
        : src(in + row * cols_), dst(out + row * cols_), cols(cols_) {
        if (WIN) {
            const unsigned mis = (unsigned)((mis0 + row * cols_) & 3u);
            head = (4u - mis) & 3u;
            if ((size_t)head > cols) head = (unsigned)cols;
            nvec = (cols - head) >> 2;
            tail = (unsigned)((cols - head) & 3u);
        } else {
            head = tail = 0;
            nvec = cols >> 2;
        }
        vsrc = reinterpret_cast<const float4*>(src + head);
        vdst = reinterpret_cast<float4*>(dst + head);
    }
    // edge element e (0 .. head + tail - 1) of the row; `fill` for every other e
    __device__ __forceinline__ float edge_load(unsigned e, float fill) const {
        if (!WIN) return fill;
        float r = fill;
        if (e < head) r = src[e];
        else if (e < head + tail) r = src[cols - tail + (e - head)];
        return r;
    }
    __device__ __forceinline__ void edge_store(unsigned e, float y) const {
        if (!WIN) return;
        if (e < head) dst[e] = y;
        else if (e < head + tail) dst[cols - tail + (e - head)] = y;
    }
};

template <bool LOG>
__device__ __forceinline__ float4 normalise4(const float4& x, float m, float lse, float inv) {
    float4 y;
    if (LOG) {  // (x - max) - ln(sum), evaluated in that order (src/vector.rs:1617-1621); x holds the inputs
        y.x = (x.x - m) - lse; y.y = (x.y - m) - lse; y.z = (x.z - m) - lse; y.w = (x.w - m) - lse;
    } else {    // x holds the exponentials
        y.x = x.x * inv; y.y = x.y * inv; y.z = x.z * inv; y.w = x.w * inv;
    }
    return y;
}

// ---- one row per CTA, row in registers, flat grid -------------------------------------------------------------
// T threads x VPT float4 hold the row (T in {256, 512}); grid = rows.  The block scheduler keeps the SM's
// load queue full across rows (several CTAs per SM), statistics are two fixed block trees.
template <int T>
__device__ __forceinline__ float block_max_t(float v, float* s_w) {
    v = warp_max(v);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = s_w[0];
#pragma unroll
    for (int w = 1; w < T / 32; ++w) r = fmaxf(r, s_w[w]);
    __syncthreads();
    return r;
}
template <int T>
__device__ __forceinline__ float block_sum_t(float v, float* s_w) {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = s_w[0];
#pragma unroll
    for (int w = 1; w < T / 32; ++w) r += s_w[w];
    __syncthreads();
    return r;
}

template <int T, int VPT, bool LOG, bool WIN>
__device__ __forceinline__ void softmax_rows_cta_body(const float* __restrict__ in, float* __restrict__ out, size_t rows,
                                                      size_t cols, unsigned mis0, float* s_w) {
    const size_t row = blockIdx.x;
    const RowView<WIN> rv(in, out, row, cols, mis0);
    const unsigned nvec = (unsigned)rv.nvec;
    float4 x[VPT];
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        const unsigned v = j * T + threadIdx.x;
        x[j] = v < nvec ? ld_stream(rv.vsrc + v) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    }
    float xe = rv.edge_load(threadIdx.x, -INFINITY);
    float m = xe;
#pragma unroll
    for (int j = 0; j < VPT; ++j) m = fmaxf(m, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
    m = block_max_t<T>(m, s_w);
    float part = 0.f;
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        float4 e;
        e.x = expf(x[j].x - m); e.y = expf(x[j].y - m); e.z = expf(x[j].z - m); e.w = expf(x[j].w - m);
        part += (e.x + e.y) + (e.z + e.w);
        if (!LOG) x[j] = e;
    }
    if (WIN) {
        const float ee = expf(xe - m);
        part += ee;
        if (!LOG) xe = ee;
    }
    const float sum = block_sum_t<T>(part, s_w);
    const float lse = LOG ? logf(sum) : 0.f;
    const float inv = LOG ? 0.f : __frcp_rn(sum);
#pragma unroll
    for (int j = 0; j < VPT; ++j) {
        const unsigned v = j * T + threadIdx.x;
        if (v < nvec) st_stream(rv.vdst + v, normalise4<LOG>(x[j], m, lse, inv));
    }
    if (WIN) rv.edge_store(threadIdx.x, LOG ? (xe - m) - lse : xe * inv);
}

// Two entry points over the same body.  The log + window variants at VPT = 8 take 67-75 registers and would halve the
// occupancy, so they are compiled with a 64-register cap (1024 / T resident CTAs); every other variant is compiled
// WITHOUT a minimum-blocks argument: naming one (even 1) changes ptxas' scheduling — 95 registers, or the same count
// and 15 % less throughput at 4099 columns (measured, scripts/sweep_rows_ext.py).
template <int T, int VPT, bool LOG, bool WIN>
__global__ void __launch_bounds__(T)
softmax_rows_cta_kernel(const float* __restrict__ in, float* __restrict__ out, size_t rows, size_t cols, unsigned mis0) {
    __shared__ float s_w[T / 32];
    softmax_rows_cta_body<T, VPT, LOG, WIN>(in, out, rows, cols, mis0, s_w);
}
template <int T, int VPT, bool LOG, bool WIN>
__global__ void __launch_bounds__(T, 1024 / T)
softmax_rows_cta_capped_kernel(const float* __restrict__ in, float* __restrict__ out, size_t rows, size_t cols, unsigned mis0) {
    __shared__ float s_w[T / 32];
    softmax_rows_cta_body<T, VPT, LOG, WIN>(in, out, rows, cols, mis0, s_w);
}

template <int T, int VPT, bool LOG, bool WIN>
static int launch_cta(const float* a, float* out, size_t rows, size_t cols, unsigned mis0, cudaStream_t s) {
    if (rows > 0x7FFFFFFFull) return fail(TRN_INVALID_INPUT, "%zu rows exceed the launch grid", rows);
    if constexpr (VPT == 8 && LOG && WIN) softmax_rows_cta_capped_kernel<T, VPT, LOG, WIN><<<(unsigned)rows, T, 0, s>>>(a, out, rows, cols, mis0);
    else                        softmax_rows_cta_kernel<T, VPT, LOG, WIN><<<(unsigned)rows, T, 0, s>>>(a, out, rows, cols, mis0);
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

// ---- TMA ring kernel -----------------------------------------------------------------------------------
static int env_int_default(const char* name, int dflt) {   // experiment knobs, read per call
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
namespace ring {
constexpr int kConsumers = 512;                 // 16 consumer warps
constexpr int kRingThreads = kConsumers + 32;   // + 1 producer warp
constexpr int kRowVec = 16 * kConsumers;        // a row is <= 16 float4 per consumer thread = 32 768 floats
constexpr uint32_t kRingBytes = 224 * 1024;     // the ring: 28 / HPC slots of HPC * 8 KiB (HPC float4 per consumer thread per slot)
constexpr int kRowQueue = 8;                    // row numbers handed from the producer to the consumers (it runs < 3 rows ahead)
constexpr uint32_t kNoRow = 0xFFFFFFFFu;
constexpr uint32_t kSmemBytes = kRingBytes + 2 * 28 * 8 + 2 * 16 * 4 + kRowQueue * 4 + 128;
}  // namespace ring

// WIN: the producer bulk-copies the aligned body of each row; the (up to six) edge elements are read from global
// memory by consumer threads 0..5.
// Rows are CLAIMED, not dealt: after its first row (blockIdx.x) a CTA's producer takes the next unclaimed row from a counter
// in the stream's workspace (`claim[0]`, offset by the grid size; `claim[1]` counts CTAs that are done, the last one zeroes
// both for the next launch) and hands the row number to the consumers through a small queue in shared memory.  The SMs do
// not stream at one pace; with dealt rows the slowest SM ends the kernel microseconds after the median one, with claimed rows
// the SMs end together.  A row's result does not depend on which CTA computes it, so nothing changes bit-wise.
// claim == nullptr: rows dealt round-robin (blockIdx.x + i * gridDim.x).
template <bool LOG, bool WIN, int HPC>
__global__ void __launch_bounds__(ring::kRingThreads, 1)
softmax_rows_ring_kernel(const float* __restrict__ in, float* __restrict__ out, size_t rows, size_t cols, unsigned mis0,
                         unsigned* __restrict__ claim) {
    using namespace ring;
    constexpr int kChunkVec = HPC * kConsumers;
    constexpr uint32_t kChunkBytes = kChunkVec * 16;
    constexpr int kMaxChunks = 16 / HPC;
    constexpr int kSlots = 28 / HPC;
    extern __shared__ uint8_t ring_smem_raw[];
    const uint32_t base = (smem_u32(ring_smem_raw) + 127u) & ~127u;
    uint8_t* gen = ring_smem_raw + (base - smem_u32(ring_smem_raw));
    const uint32_t bar_base = base + kSlots * kChunkBytes;
    auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (kSlots + s); };
    float* s_max = reinterpret_cast<float*>(gen + kSlots * kChunkBytes + 2 * kSlots * 8);
    float* s_sum = s_max + 16;
    volatile uint32_t* s_rowq = reinterpret_cast<volatile uint32_t*>(s_sum + 16);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < (uint32_t)kSlots; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), kConsumers / 32); }
        fence_barrier_init();
    }
    __syncthreads();

    if (warp == kConsumers / 32) {
        // ===================== producer: one elected lane streams rows into the ring =====================
        if (elect_one()) {
            uint32_t slot = 0, phase = 0, seq = 0;
            size_t row = blockIdx.x;
            for (;; ++seq) {
                if (row >= rows) {
                    // no row left: a queue entry that says so, published by completing the next slot's phase without data
                    s_rowq[seq % kRowQueue] = kNoRow;
                    mbar_wait(empty_bar(slot), phase ^ 1);
                    mbar_arrive(full_bar(slot));
                    break;
                }
                s_rowq[seq % kRowQueue] = (uint32_t)row;     // ordered before the row's first full-barrier arrive (release)
                const RowView<WIN> rv(in, out, row, cols, mis0);
                const unsigned nvec = (unsigned)rv.nvec;
                const unsigned nchunks = (nvec + kChunkVec - 1) / kChunkVec;
                const uint32_t row_bytes = nvec * 16u;
                const char* src = reinterpret_cast<const char*>(rv.vsrc);
                for (unsigned j = 0; j < nchunks; ++j) {
                    mbar_wait(empty_bar(slot), phase ^ 1);
                    const uint32_t off = j * kChunkBytes;
                    const uint32_t bytes = row_bytes - off < kChunkBytes ? row_bytes - off : kChunkBytes;
                    mbar_expect_tx(full_bar(slot), bytes);
                    bulk_load_1d(base + slot * kChunkBytes, src + off, bytes, full_bar(slot));
                    if (++slot == (uint32_t)kSlots) { slot = 0; phase ^= 1; }
                }
                // the next row: claimed while this one is still landing (the atomic's round trip hides under the ring)
                row = claim ? (size_t)gridDim.x + atomicAdd(&claim[0], 1u) : row + gridDim.x;
            }
            if (claim && atomicAdd(&claim[1], 1u) == gridDim.x - 1) {   // every CTA has made its last claim: reset for the next launch
                claim[0] = 0;
                claim[1] = 0;
            }
        }
        return;
    }

    // ===================== consumers =====================
    const int t = threadIdx.x;
    uint32_t slot = 0, phase = 0;
    for (uint32_t seq = 0;; ++seq) {
        mbar_wait(full_bar(slot), phase);              // the next row's first chunk, or the "no row left" entry
        const uint32_t qrow = s_rowq[seq % kRowQueue];
        if (qrow == kNoRow) break;
        const size_t row = qrow;
        const RowView<WIN> rv(in, out, row, cols, mis0);
        const unsigned nvec = (unsigned)rv.nvec;
        const unsigned nchunks = (nvec + kChunkVec - 1) / kChunkVec;
        float xe = rv.edge_load((unsigned)t, -INFINITY);
        float4 x[HPC * kMaxChunks];
        // ---- ring -> registers; each slot goes back to the producer as soon as it has been read
#pragma unroll
        for (int j = 0; j < kMaxChunks; ++j) {
            if ((unsigned)j < nchunks) {
                mbar_wait(full_bar(slot), phase);
                const uint32_t sb = base + slot * kChunkBytes;
#pragma unroll
                for (int h = 0; h < HPC; ++h) {
                    const unsigned v = j * kChunkVec + h * kConsumers + t;
                    float4 r = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                    if (v < nvec)
                        asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                     : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(sb + (h * kConsumers + t) * 16u));
                    x[HPC * j + h] = r;
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty_bar(slot));
                if (++slot == (uint32_t)kSlots) { slot = 0; phase ^= 1; }
            } else {
#pragma unroll
                for (int h = 0; h < HPC; ++h) x[HPC * j + h] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
            }
        }

        // ---- row max: warp tree, then a fixed-order fold of the 16 warp values in every thread
        float m = xe;
#pragma unroll
        for (int j = 0; j < HPC * kMaxChunks; ++j) m = fmaxf(m, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
        m = warp_max(m);
        if (lane == 0) s_max[warp] = m;
        named_bar_sync(1, kConsumers);
        m = s_max[0];
#pragma unroll
        for (int w = 1; w < kConsumers / 32; ++w) m = fmaxf(m, s_max[w]);

        // ---- exponentials and their sum.  Padding slots hold -inf -> expf(-inf) = 0 exactly.  The sum is a TREE at every
        // level (four interleaved chains per thread, shuffle tree, pairwise fold of the warp sums): the row sum's rounding
        // error lands on every element of the row, and a 16-long chain per thread plus a 16-long chain over the warps put the
        // worst element of a 16 M-element input at 9.5 ulp where the 8-ulp contract allows expf's own 2 ulp little company.
        float part4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int j = 0; j < HPC * kMaxChunks; ++j) {
            float4 e;
            e.x = expf(x[j].x - m); e.y = expf(x[j].y - m); e.z = expf(x[j].z - m); e.w = expf(x[j].w - m);
            part4[j & 3] += (e.x + e.y) + (e.z + e.w);
            if (!LOG) x[j] = e;
        }
        float part = (part4[0] + part4[1]) + (part4[2] + part4[3]);
        if (WIN) {
            const float ee = expf(xe - m);
            part += ee;
            if (!LOG) xe = ee;
        }
        part = warp_sum(part);
        if (lane == 0) s_sum[warp] = part;
        named_bar_sync(1, kConsumers);
        float ws[kConsumers / 32];
#pragma unroll
        for (int w = 0; w < kConsumers / 32; ++w) ws[w] = s_sum[w];
#pragma unroll
        for (int step = 1; step < kConsumers / 32; step *= 2) {
#pragma unroll
            for (int w = 0; w + step < kConsumers / 32; w += 2 * step) ws[w] += ws[w + step];
        }
        const float sum = ws[0];
        // (after THIS barrier every thread has read s_max; s_sum is rewritten after the next row's max barrier.)

        // ---- scale and store straight from registers (512 contiguous bytes per warp per store)
        const float lse = LOG ? logf(sum) : 0.f;
        const float inv = LOG ? 0.f : __frcp_rn(sum);
#pragma unroll
        for (int j = 0; j < HPC * kMaxChunks; ++j) {
            const unsigned v = j * kConsumers + t;   // = (j / HPC) * kChunkVec + (j % HPC) * kConsumers + t
            if (v < nvec) st_stream(rv.vdst + v, normalise4<LOG>(x[j], m, lse, inv));
        }
        if (WIN) rv.edge_store((unsigned)t, LOG ? (xe - m) - lse : xe * inv);
    }
}

template <bool LOG, bool WIN, int HPC>
static int launch_ring(const float* a, float* out, size_t rows, size_t cols, unsigned mis0, int sm_count, cudaStream_t s) {
    static const cudaError_t attr = cudaFuncSetAttribute(softmax_rows_ring_kernel<LOG, WIN, HPC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                         (int)ring::kSmemBytes);   // thread-safe, once
    TRN_CUDA(attr);
    // Dealt rows (TRN_RING_DYN=0, or no workspace) go out in even waves: every persistent CTA walks ceil(rows / SMs) rows, so
    // the grid is the smallest one that still does — 512 rows (4096 rows sharded over 8 GPUs) run as 128 CTAs x 4 rows instead
    // of 148 CTAs of which 68 take a 4th row while 80 idle; a single SM's ring can absorb the bandwidth the idle ones leave
    // (~60 vs 42 GB/s per SM).
    const size_t waves = (rows + (size_t)sm_count - 1) / (size_t)sm_count;
    unsigned grid = (unsigned)(waves ? (rows + waves - 1) / waves : 1);
    // claimed rows (default; TRN_RING_DYN=0 deals them): one CTA per SM, the counters live behind the stream's reduction ticket
    unsigned* claim = nullptr;
    if (env_int_default("TRN_RING_DYN", 1) && rows < 0xFFFFFFFFull - 1024) {
        Workspace* w = workspace(s);
        if (w) {
            claim = w->ticket + 4;
            grid = (unsigned)(rows < (size_t)sm_count ? rows : (size_t)sm_count);
        }
    }
    softmax_rows_ring_kernel<LOG, WIN, HPC><<<grid, ring::kRingThreads, ring::kSmemBytes, s>>>(a, out, rows, cols, mis0, claim);
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

// ---- config-5 rows over CTA PAIRS: two half-rows in flight per SM ------------------------------------------------------
// The ring kernel above keeps ONE row per SM in flight and all sixteen consumer warps walk its phases in lock step (shared
// memory -> registers, max, exp, sum, store): while they compute nothing is stored, while they store nothing is computed,
// and a launch pays a whole 128 KB row of ramp-up and of drain per SM.  Here a row belongs to a 2-CTA cluster: each CTA
// streams ITS HALF of every row of the cluster through its own ring, and its consumer warps form two groups of eight that
// take alternate rows — two half-rows per SM in different phases, half the ramp, half the wave quantum (512 rows on 148 SMs:
// 6.92 half-rows per SM instead of 3.46 rows).  The two halves of a row meet ONCE: each group folds its half to an online
// (max, sum of exp relative to it) pair and sends it to the partner CTA with one `st.async` (8 bytes + the partner's
// mbarrier transaction count: no fence, no cluster barrier in the row loop); both CTAs fold the two pairs in rank order,
// so they normalise by identical bits.  softmax = exp(x - m_half) * (exp(m_half - M) / S), log_softmax = (x - M) - ln S.
namespace ring2 {
constexpr int kGroup = 256;                        // consumer threads per group
constexpr int kThreads2 = 2 * kGroup + 32;         // two groups + the producer warp
constexpr int kSlotVec = 4 * kGroup;               // float4 per slot: 4 per consumer thread (16 KiB)
constexpr uint32_t kSlotBytes = kSlotVec * 16;
constexpr int kSlots = 14;                         // 224 KiB of ring
constexpr int kMaxChunks = 4;                      // a half-row is <= 16 float4 per thread = 16 384 floats
constexpr uint32_t kSmemBytes = kSlots * kSlotBytes + 2 * kSlots * 8 + 4 * 8 /*xbar*/ + 4 * 8 /*mailbox*/ + 2 * 2 * 8 * 4 /*s_max, s_sum*/ + 128;
}  // namespace ring2

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta_rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta_rank));
    return r;
}

template <bool LOG>
__global__ void __launch_bounds__(ring2::kThreads2, 1)
softmax_rows_ring2_kernel(const float* __restrict__ in, float* __restrict__ out, size_t rows, size_t cols) {
    using namespace ring2;
    extern __shared__ uint8_t ring2_smem_raw[];
    const uint32_t base = (smem_u32(ring2_smem_raw) + 127u) & ~127u;
    uint8_t* gen = ring2_smem_raw + (base - smem_u32(ring2_smem_raw));
    const uint32_t bar_base = base + kSlots * kSlotBytes;
    auto full_bar = [&](uint32_t s) { return bar_base + 8u * s; };
    auto empty_bar = [&](uint32_t s) { return bar_base + 8u * (kSlots + s); };
    auto xbar = [&](uint32_t g, uint32_t par) { return bar_base + 8u * (2 * kSlots + 2 * g + par); };
    const uint32_t mbox_base = bar_base + 8u * (2 * kSlots + 4);
    auto mbox = [&](uint32_t g, uint32_t par) { return mbox_base + 8u * (2 * g + par); };
    uint8_t* tail = gen + kSlots * kSlotBytes + 8 * (2 * kSlots + 4);
    const volatile float2* mbox_gen = reinterpret_cast<const volatile float2*>(tail);
    float* s_max = reinterpret_cast<float*>(tail + 4 * 8);   // [2 groups][8 warps]
    float* s_sum = s_max + 16;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const size_t cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const size_t my_rows = rows > cluster_id ? (rows - cluster_id + num_clusters - 1) / num_clusters : 0;
    const unsigned hv = (unsigned)(cols >> 3);                 // float4 per half-row
    const unsigned nchunks = (hv + kSlotVec - 1) / kSlotVec;

    if (threadIdx.x == 0) {
        for (uint32_t s = 0; s < (uint32_t)kSlots; ++s) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), kGroup / 32); }
        for (uint32_t g = 0; g < 2; ++g) for (uint32_t par = 0; par < 2; ++par) mbar_init(xbar(g, par), 1);
        fence_barrier_init();
    }
    // the partner's messages must not arrive before these barriers exist
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");

    if (warp == 2 * kGroup / 32) {
        // ===================== producer: this CTA's half of every row of the cluster, in row order =====================
        if (elect_one()) {
            uint32_t slot = 0, phase = 0;
            for (size_t i = 0; i < my_rows; ++i) {
                const size_t row = cluster_id + i * num_clusters;
                const char* src = reinterpret_cast<const char*>(in + row * cols) + (size_t)rank * hv * 16u;
                const uint32_t half_bytes = hv * 16u;
                for (unsigned j = 0; j < nchunks; ++j) {
                    mbar_wait(empty_bar(slot), phase ^ 1);
                    const uint32_t off = j * kSlotBytes;
                    const uint32_t bytes = half_bytes - off < kSlotBytes ? half_bytes - off : kSlotBytes;
                    mbar_expect_tx(full_bar(slot), bytes);
                    bulk_load_1d(base + slot * kSlotBytes, src + off, bytes, full_bar(slot));
                    if (++slot == (uint32_t)kSlots) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===================== consumers: group g takes the cluster's rows g, g + 2, g + 4, ... =====================
        const uint32_t g = (uint32_t)warp >> 3;
        const int t = (int)threadIdx.x - (int)g * kGroup;
        const int gw = warp & 7;
        float* gmax = s_max + g * 8;
        float* gsum = s_sum + g * 8;
        const uint32_t partner = rank ^ 1u;
        const uint32_t n_mine = (uint32_t)my_rows;
        for (uint32_t i = g, k = 0; i < n_mine; i += 2, ++k) {
            const size_t row = cluster_id + (size_t)i * num_clusters;
            float4 x[4 * kMaxChunks];
            // ---- ring -> registers; each slot goes back to the producer as soon as this group has read it
#pragma unroll
            for (int j = 0; j < kMaxChunks; ++j) {
                if ((unsigned)j < nchunks) {
                    const uint32_t cnt = i * nchunks + (uint32_t)j;      // position of this chunk in the producer's order
                    const uint32_t slot = cnt % (uint32_t)kSlots, phase = (cnt / (uint32_t)kSlots) & 1u;
                    mbar_wait(full_bar(slot), phase);
                    const uint32_t sb = base + slot * kSlotBytes;
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        const unsigned v = j * kSlotVec + h * kGroup + t;
                        float4 r = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                        if (v < hv)
                            asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];"
                                         : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(sb + (h * kGroup + t) * 16u));
                        x[4 * j + h] = r;
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(empty_bar(slot));
                } else {
#pragma unroll
                    for (int h = 0; h < 4; ++h) x[4 * j + h] = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
                }
            }
            // ---- this half's max: warp tree, then a fixed-order fold of the group's 8 warp values in every thread
            float m = -INFINITY;
#pragma unroll
            for (int j = 0; j < 4 * kMaxChunks; ++j) m = fmaxf(m, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
            m = warp_max(m);
            if (lane == 0) gmax[gw] = m;
            named_bar_sync(1 + g, kGroup);
            m = gmax[0];
#pragma unroll
            for (int w = 1; w < 8; ++w) m = fmaxf(m, gmax[w]);
            // ---- exponentials relative to THIS half's max and their sum (padding holds -inf -> exactly 0).  A half that is
            // all -inf must contribute (max -inf, sum 0), not exp(-inf - -inf) = NaN: its exponentials are taken relative to 0
            // (the row is NaN only when BOTH halves are -inf, through exp(m - M) below, as the reference's whole-row form is)
            const float mref = m == -INFINITY ? 0.f : m;
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < 4 * kMaxChunks; ++j) {
                float4 e;
                e.x = expf(x[j].x - mref); e.y = expf(x[j].y - mref); e.z = expf(x[j].z - mref); e.w = expf(x[j].w - mref);
                part += (e.x + e.y) + (e.z + e.w);
                if (!LOG) x[j] = e;
            }
            part = warp_sum(part);
            if (lane == 0) gsum[gw] = part;
            named_bar_sync(1 + g, kGroup);
            float sum = gsum[0];
#pragma unroll
            for (int w = 1; w < 8; ++w) sum += gsum[w];
            // ---- one exchange with the other half of the row: (m, sum) in one 8-byte st.async to the partner's mailbox
            const uint32_t par = k & 1u, xphase = (k >> 1) & 1u;
            if (t == 0) {
                mbar_expect_tx(xbar(g, par), 8);
                const unsigned long long msg = ((unsigned long long)__float_as_uint(sum) << 32) | __float_as_uint(m);
                asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
                             :: "r"(mapa_u32(mbox(g, par), partner)), "l"(msg), "r"(mapa_u32(xbar(g, par), partner)) : "memory");
            }
            mbar_wait(xbar(g, par), xphase);
            const float2 other = make_float2(mbox_gen[2 * g + par].x, mbox_gen[2 * g + par].y);   // {max, sum} of the other half
            // fold in rank order: both CTAs of the pair get the same bits
            const float m0 = rank == 0 ? m : other.x, s0 = rank == 0 ? sum : other.y;
            const float m1 = rank == 0 ? other.x : m, s1 = rank == 0 ? other.y : sum;
            const float M = fmaxf(m0, m1);
            const float S = s0 * expf(m0 - M) + s1 * expf(m1 - M);
            const float lse = LOG ? logf(S) : 0.f;
            const float f = LOG ? 0.f : expf(m - M) / S;
            float4* dst = reinterpret_cast<float4*>(out + row * cols) + (size_t)rank * hv;
#pragma unroll
            for (int j = 0; j < 4 * kMaxChunks; ++j) {
                const unsigned v = j * kGroup + t;   // = (j / 4) * kSlotVec + (j % 4) * kGroup + t
                if (v < hv) st_stream(dst + v, normalise4<LOG>(x[j], M, lse, f));
            }
            // every thread of the group has read the mailbox before its leader can announce the next row on this parity
            named_bar_sync(1 + g, kGroup);
        }
    }
    // neither CTA may leave while its partner can still write into its shared memory
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <bool LOG>
static int launch_ring2(const float* a, float* out, size_t rows, size_t cols, int sm_count, cudaStream_t s) {
    static const cudaError_t attr = cudaFuncSetAttribute(softmax_rows_ring2_kernel<LOG>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                         (int)ring2::kSmemBytes);   // thread-safe, once
    TRN_CUDA(attr);
    const size_t max_clusters = (size_t)sm_count / 2;
    const size_t waves = (rows + max_clusters - 1) / max_clusters;       // even waves, as for the ring kernel
    const size_t clusters = waves ? (rows + waves - 1) / waves : 1;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(ring2::kThreads2);
    cfg.dynamicSmemBytes = ring2::kSmemBytes;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    TRN_CUDA(cudaLaunchKernelEx(&cfg, softmax_rows_ring2_kernel<LOG>, a, out, rows, cols));
    count_launch();
    return TRN_OK;
}

// ---- short rows: one WARP per row -------------------------------------------------------------------------
// cols <= 1024 (attention-sized rows): a row is VPT float4 per lane, RPW rows per warp are loaded up front
// (RPW * VPT = 8 independent 128-bit loads in flight per lane), all statistics are warp shuffles — no block
// barrier at all — and a CTA of 8 warps moves 8 * RPW rows.  Flat grid.  MODE 0 softmax, 1 log_softmax,
// 2 layer_norm (gamma / beta shared by all rows, src/vector.rs:1316-1362; both null = layer_norm_simple, :1386).
template <int VPT, int MODE, bool WIN>
__global__ void __launch_bounds__(kThreads)
rows_warp_kernel(const float* __restrict__ in, float* __restrict__ out, size_t rows, size_t cols,
                 const float* __restrict__ gamma, const float* __restrict__ beta, float eps, unsigned mis0) {
    static_assert(!(WIN && MODE == 2), "layer_norm has no window form");
    constexpr int RPW = 8 / VPT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t row0 = ((size_t)blockIdx.x * (kThreads / 32) + warp) * RPW;
    const float fill = MODE == 2 ? 0.f : -INFINITY;
    float4 x[RPW][VPT];
    float xe[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const size_t row = row0 + r;
        const RowView<WIN> rv(in, out, row < rows ? row : 0, cols, mis0);
        const unsigned nvec = (unsigned)rv.nvec;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const unsigned v = j * 32 + lane;
            x[r][j] = (row < rows && v < nvec) ? ld_stream(rv.vsrc + v) : make_float4(fill, fill, fill, fill);
        }
        xe[r] = (WIN && row < rows) ? rv.edge_load((unsigned)lane, fill) : fill;
    }
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
        const size_t row = row0 + r;
        if (row >= rows) break;   // warp-uniform
        const RowView<WIN> rv(in, out, row, cols, mis0);
        const unsigned nvec = (unsigned)rv.nvec;
        if (MODE == 2) {
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < VPT; ++j) part += (x[r][j].x + x[r][j].y) + (x[r][j].z + x[r][j].w);
            const float mean = warp_sum(part) / (float)cols;
            part = 0.f;
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                if (j * 32 + lane < (int)nvec) {
                    const float a = __fsub_rn(x[r][j].x, mean), b = __fsub_rn(x[r][j].y, mean);
                    const float c = __fsub_rn(x[r][j].z, mean), d = __fsub_rn(x[r][j].w, mean);
                    part += (__fmul_rn(a, a) + __fmul_rn(b, b)) + (__fmul_rn(c, c) + __fmul_rn(d, d));
                }
            }
            const float inv_std = 1.0f / sqrtf(warp_sum(part) / (float)cols + eps);
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const unsigned v = j * 32 + lane;
                if (v < nvec) {
                    float4 y;
                    if (gamma) {   // gamma * (x - mean) * inv_std + beta, evaluated in that order, unfused
                        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + v);
                        const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + v);
                        y.x = __fadd_rn(__fmul_rn(__fmul_rn(g.x, __fsub_rn(x[r][j].x, mean)), inv_std), bt.x);
                        y.y = __fadd_rn(__fmul_rn(__fmul_rn(g.y, __fsub_rn(x[r][j].y, mean)), inv_std), bt.y);
                        y.z = __fadd_rn(__fmul_rn(__fmul_rn(g.z, __fsub_rn(x[r][j].z, mean)), inv_std), bt.z);
                        y.w = __fadd_rn(__fmul_rn(__fmul_rn(g.w, __fsub_rn(x[r][j].w, mean)), inv_std), bt.w);
                    } else {       // layer_norm_simple: (x - mean) * inv_std
                        y.x = __fmul_rn(__fsub_rn(x[r][j].x, mean), inv_std);
                        y.y = __fmul_rn(__fsub_rn(x[r][j].y, mean), inv_std);
                        y.z = __fmul_rn(__fsub_rn(x[r][j].z, mean), inv_std);
                        y.w = __fmul_rn(__fsub_rn(x[r][j].w, mean), inv_std);
                    }
                    st_stream(rv.vdst + v, y);
                }
            }
        } else {
            float m = xe[r];
#pragma unroll
            for (int j = 0; j < VPT; ++j) m = fmaxf(m, fmaxf(fmaxf(x[r][j].x, x[r][j].y), fmaxf(x[r][j].z, x[r][j].w)));
            m = warp_max(m);
            float part = 0.f;
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                float4 e;
                e.x = expf(x[r][j].x - m); e.y = expf(x[r][j].y - m); e.z = expf(x[r][j].z - m); e.w = expf(x[r][j].w - m);
                part += (e.x + e.y) + (e.z + e.w);
                if (MODE == 0) x[r][j] = e;
            }
            float xer = xe[r];
            if (WIN) {
                const float ee = expf(xer - m);
                part += ee;
                if (MODE == 0) xer = ee;
            }
            const float sum = warp_sum(part);
            const float lse = MODE == 1 ? logf(sum) : 0.f;
            const float inv = MODE == 1 ? 0.f : __frcp_rn(sum);
#pragma unroll
            for (int j = 0; j < VPT; ++j) {
                const unsigned v = j * 32 + lane;
                if (v < nvec) st_stream(rv.vdst + v, normalise4<MODE == 1>(x[r][j], m, lse, inv));
            }
            if (WIN) rv.edge_store((unsigned)lane, MODE == 1 ? (xer - m) - lse : xer * inv);
        }
    }
}

template <int MODE, bool WIN>
static int launch_rows_warp(const float* a, float* out, size_t rows, size_t cols, const float* gamma, const float* beta,
                            float eps, unsigned mis0, cudaStream_t s) {
    const size_t nvec = cols / 4;   // the body of a window row is never longer
    const int vpt = nvec <= 32 ? 1 : nvec <= 64 ? 2 : nvec <= 128 ? 4 : 8;
    const size_t rows_per_cta = (size_t)(kThreads / 32) * (8 / vpt);
    const size_t grid = (rows + rows_per_cta - 1) / rows_per_cta;
    if (grid > 0x7FFFFFFFull) return fail(TRN_INVALID_INPUT, "%zu rows exceed the launch grid", rows);
    switch (vpt) {
        case 1: rows_warp_kernel<1, MODE, WIN><<<(unsigned)grid, kThreads, 0, s>>>(a, out, rows, cols, gamma, beta, eps, mis0); break;
        case 2: rows_warp_kernel<2, MODE, WIN><<<(unsigned)grid, kThreads, 0, s>>>(a, out, rows, cols, gamma, beta, eps, mis0); break;
        case 4: rows_warp_kernel<4, MODE, WIN><<<(unsigned)grid, kThreads, 0, s>>>(a, out, rows, cols, gamma, beta, eps, mis0); break;
        default: rows_warp_kernel<8, MODE, WIN><<<(unsigned)grid, kThreads, 0, s>>>(a, out, rows, cols, gamma, beta, eps, mis0); break;
    }
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

// layer_norm with the row in the registers of one CTA (1024 < cols <= 16384): one read, one write
template <int T, int VPT>
__global__ void __launch_bounds__(T)
layer_norm_rows_reg_kernel(const float* __restrict__ in, const float* __restrict__ gamma, const float* __restrict__ beta,
                           float eps, float* __restrict__ out, size_t rows, size_t cols) {
    __shared__ float s_w[T / 32];
    const unsigned nvec = (unsigned)(cols >> 2);
    for (size_t row = blockIdx.x; row < rows; row += gridDim.x) {
        const float4* src = reinterpret_cast<const float4*>(in + row * cols);
        float4* dst = reinterpret_cast<float4*>(out + row * cols);
        float4 x[VPT];
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const unsigned v = j * T + threadIdx.x;
            x[j] = v < nvec ? ld_stream(src + v) : make_float4(0, 0, 0, 0);
        }
        float part = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) part += (x[j].x + x[j].y) + (x[j].z + x[j].w);
        const float mean = block_sum_t<T>(part, s_w) / (float)cols;
        part = 0.f;
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            if (j * T + threadIdx.x < nvec) {
                const float a = __fsub_rn(x[j].x, mean), b = __fsub_rn(x[j].y, mean);
                const float c = __fsub_rn(x[j].z, mean), d = __fsub_rn(x[j].w, mean);
                part += (__fmul_rn(a, a) + __fmul_rn(b, b)) + (__fmul_rn(c, c) + __fmul_rn(d, d));
            }
        }
        const float inv_std = 1.0f / sqrtf(block_sum_t<T>(part, s_w) / (float)cols + eps);
#pragma unroll
        for (int j = 0; j < VPT; ++j) {
            const unsigned v = j * T + threadIdx.x;
            if (v < nvec) {
                float4 y;
                if (gamma) {
                    const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + v);
                    const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + v);
                    y.x = __fadd_rn(__fmul_rn(__fmul_rn(g.x, __fsub_rn(x[j].x, mean)), inv_std), bt.x);
                    y.y = __fadd_rn(__fmul_rn(__fmul_rn(g.y, __fsub_rn(x[j].y, mean)), inv_std), bt.y);
                    y.z = __fadd_rn(__fmul_rn(__fmul_rn(g.z, __fsub_rn(x[j].z, mean)), inv_std), bt.z);
                    y.w = __fadd_rn(__fmul_rn(__fmul_rn(g.w, __fsub_rn(x[j].w, mean)), inv_std), bt.w);
                } else {
                    y.x = __fmul_rn(__fsub_rn(x[j].x, mean), inv_std);
                    y.y = __fmul_rn(__fsub_rn(x[j].y, mean), inv_std);
                    y.z = __fmul_rn(__fsub_rn(x[j].z, mean), inv_std);
                    y.w = __fmul_rn(__fsub_rn(x[j].w, mean), inv_std);
                }
                st_stream(dst + v, y);
            }
        }
    }
}

// ---- LONG rows: a row over a CTA cluster, two passes, online (max, sum) -------------------------------------
// cols > 32 768.  Pass 1: every thread walks its share of the row in steps of U independent 128-bit loads and keeps
// a running (m, s) pair, s = sum of exp(x - m) over what it has seen; a step costs one rescale exp(m_old - m_new)
// per 16 elements.  The pairs fold to one (M, S) per CTA through a fixed block tree and to the row's (M, S) through
// distributed shared memory in rank order -> deterministic.  Pass 2 re-reads the row (L2: a row is a few MB at most
// against 126 MB) and writes exp(x - M) / S or (x - M) - ln S, the reference's expressions
// (src/vector.rs:1540-1553, :1605-1623).  An all -inf prefix keeps s = 0 (reference: exp(-inf - max) = 0); an
// all -inf ROW gives NaN as the reference does (x - max = -inf - -inf).
// Measured against two one-pass forms, both since removed (scripts/exp/exp_long_rows.py): a register-resident 8-CTA
// cluster kernel (40 000 columns 5.3 vs 3.3 TB/s, 65 536: 5.2 vs 4.7) and a shared-memory-resident cluster kernel (every
// CTA bulk-copies its slice of the row into shared memory, one exponential per element, pass 2 from shared memory:
// 50 257 columns 5.1 vs 4.3, 128 256: 5.0 vs 4.5 — one row per CTA leaves the TMA round trip and the cluster barrier
// uncovered, where this kernel keeps eight CTAs of plain loads per SM in flight).
__device__ __forceinline__ void online_merge(float& m, float& s, float m2, float s2) {
    const float mn = fmaxf(m, m2);
    const float ref = mn == -INFINITY ? 0.f : mn;   // nothing finite seen yet: every term is exp(-inf) = 0
    s = s * expf(m - ref) + s2 * expf(m2 - ref);
    m = mn;
}

// online (max, sum) over body vectors [v0, v1) taken kThreads * U at a time with stride `stride` between a thread's
// U vectors (callers interleave CTAs of a cluster through `first` / `stride`)
template <int U>
__device__ __forceinline__ void online_step(const float4 (&x)[U], float& m, float& s) {
    float tm = -INFINITY;
#pragma unroll
    for (int j = 0; j < U; ++j) tm = fmaxf(tm, fmaxf(fmaxf(x[j].x, x[j].y), fmaxf(x[j].z, x[j].w)));
    const float mn = fmaxf(m, tm);
    const float ref = mn == -INFINITY ? 0.f : mn;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < U; ++j)
        acc += (expf(x[j].x - ref) + expf(x[j].y - ref)) + (expf(x[j].z - ref) + expf(x[j].w - ref));
    s = s * expf(m - ref) + acc;
    m = mn;
}

template <bool LOG>
__device__ __forceinline__ float4 normalise_raw4(const float4& x, float M, float lse, float inv) {
    float4 y;
    if (LOG) {
        y.x = (x.x - M) - lse; y.y = (x.y - M) - lse; y.z = (x.z - M) - lse; y.w = (x.w - M) - lse;
    } else {
        y.x = expf(x.x - M) * inv; y.y = expf(x.y - M) * inv; y.z = expf(x.z - M) * inv; y.w = expf(x.w - M) * inv;
    }
    return y;
}

template <int CS, bool LOG, bool WIN, bool HINT>
__global__ void __launch_bounds__(kThreads)
softmax_rows_long_kernel(const float* __restrict__ in, float* __restrict__ out, size_t rows, size_t cols, unsigned mis0) {
    constexpr int U = 4;
    __shared__ float s_w[kThreads / 32];
    __shared__ float s_stat[2];  // [0] CTA max, [1] CTA exp-sum relative to it — read by cluster peers through DSMEM

    unsigned rank = 0;
    if (CS > 1) rank = cg::this_cluster().block_rank();
    const size_t cluster_id = blockIdx.x / CS;
    const size_t num_clusters = gridDim.x / CS;
    const float4 ninf4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    const uint64_t keep = l2_policy_keep(), drop = l2_policy_drop();   // pass 1 parks the row in L2, pass 2 releases it

    for (size_t row = cluster_id; row < rows; row += num_clusters) {
        const RowView<WIN> rv(in, out, row, cols, mis0);
        const size_t nvec = rv.nvec;
        const size_t step = (size_t)CS * kThreads * U;   // vectors the cluster consumes per step

        // ---- pass 1: online (max, sum); the edge elements ride in rank 0's threads 0..5
        const float xe = (WIN && rank == 0) ? rv.edge_load(threadIdx.x, -INFINITY) : -INFINITY;
        float m = -INFINITY, s = 0.f;
        for (size_t base = 0; base < nvec; base += step) {
            float4 x[U];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const size_t v = base + ((size_t)j * CS + rank) * kThreads + threadIdx.x;
                x[j] = v < nvec ? (HINT ? ld_stream_hint(rv.vsrc + v, keep) : ld_stream(rv.vsrc + v)) : ninf4;
            }
            online_step<U>(x, m, s);
        }
        if (WIN) online_merge(m, s, xe, xe == -INFINITY ? 0.f : 1.f);

        // ---- fold the pairs: block tree, then the cluster in rank order
        float M = block_max_256(m, s_w);
        {
            const float ref = M == -INFINITY ? 0.f : M;
            s = s * expf(m - ref);
        }
        float S = block_sum_256(s, s_w);
        if (CS > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            if (threadIdx.x == 0) { s_stat[0] = M; s_stat[1] = S; }
            cluster.sync();
            float gm = -INFINITY, gs = 0.f;
#pragma unroll
            for (int p = 0; p < CS; ++p)
                online_merge(gm, gs, *cluster.map_shared_rank(&s_stat[0], p), *cluster.map_shared_rank(&s_stat[1], p));
            M = gm;
            S = gs;
        }

        // ---- pass 2: re-read (L2), normalise, store
        const float lse = LOG ? logf(S) : 0.f;
        const float inv = LOG ? 0.f : __frcp_rn(S);
        for (size_t base = 0; base < nvec; base += step) {
            float4 x[U];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const size_t v = base + ((size_t)j * CS + rank) * kThreads + threadIdx.x;
                x[j] = v < nvec ? (HINT ? ld_stream_hint(rv.vsrc + v, drop) : ld_stream(rv.vsrc + v)) : ninf4;
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const size_t v = base + ((size_t)j * CS + rank) * kThreads + threadIdx.x;
                if (v < nvec) st_stream(rv.vdst + v, normalise_raw4<LOG>(x[j], M, lse, inv));
            }
        }
        if (WIN && rank == 0) rv.edge_store(threadIdx.x, LOG ? (xe - M) - lse : expf(xe - M) * inv);
        // peers must be done reading this CTA's s_stat before the next row overwrites it
        if (CS > 1) cg::this_cluster().sync();
    }
}

// ---- FEW long rows (Vector::softmax on one large vector): a row split over many CTAs, two launches ---------
// With fewer rows than SM-eighths the cluster kernel above would leave most of the machine idle (one 4 GiB vector =
// one cluster).  Here every row is cut into P segments, grid (P, rows): launch 1 folds each segment to an online
// (max, sum) pair in the workspace; launch 2 has every CTA fold its row's P pairs in a fixed tree (so all CTAs of a
// row get the same bits), then re-reads its own segment (L2 when the row fits) and writes the result.
template <bool WIN>
__global__ void __launch_bounds__(kThreads)
softmax_split_stats_kernel(const float* __restrict__ in, size_t rows, size_t cols, unsigned mis0, size_t seg,
                           float2* __restrict__ ws) {
    constexpr int U = 4;
    __shared__ float s_w[kThreads / 32];
    const size_t row = blockIdx.y;
    const RowView<WIN> rv(in, nullptr, row, cols, mis0);
    const size_t v0 = (size_t)blockIdx.x * seg;
    const size_t v1 = v0 + seg < rv.nvec ? v0 + seg : rv.nvec;
    const float4 ninf4 = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    const uint64_t keep = l2_policy_keep();   // the write launch re-reads the row: park it in L2 if it fits
    float m = -INFINITY, s = 0.f;
    for (size_t base = v0; base < v1; base += (size_t)kThreads * U) {
        float4 x[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const size_t v = base + (size_t)j * kThreads + threadIdx.x;
            x[j] = v < v1 ? ld_stream_hint(rv.vsrc + v, keep) : ninf4;
        }
        online_step<U>(x, m, s);
    }
    if (WIN && blockIdx.x == 0) {   // the edge elements ride in segment 0's threads 0..5
        const float xe = rv.edge_load(threadIdx.x, -INFINITY);
        online_merge(m, s, xe, xe == -INFINITY ? 0.f : 1.f);
    }
    const float M = block_max_256(m, s_w);
    const float ref = M == -INFINITY ? 0.f : M;
    const float S = block_sum_256(s * expf(m - ref), s_w);
    if (threadIdx.x == 0) ws[row * gridDim.x + blockIdx.x] = make_float2(M, S);
}

template <bool LOG, bool WIN>
__global__ void __launch_bounds__(kThreads)
softmax_split_write_kernel(const float* __restrict__ in, float* __restrict__ out, size_t rows, size_t cols, unsigned mis0,
                           size_t seg, const float2* __restrict__ ws, unsigned P) {
    constexpr int U = 4;
    __shared__ float s_w[kThreads / 32];
    const size_t row = blockIdx.y;
    // the row's (M, S) from its P segment pairs: thread t folds pairs t, t + 256, ... in order, then fixed block trees
    float m = -INFINITY, s = 0.f;
    for (unsigned p = threadIdx.x; p < P; p += kThreads) {
        const float2 q = ws[row * P + p];
        online_merge(m, s, q.x, q.y);
    }
    const float M = block_max_256(m, s_w);
    const float ref = M == -INFINITY ? 0.f : M;
    const float S = block_sum_256(s * expf(m - ref), s_w);

    const RowView<WIN> rv(in, out, row, cols, mis0);
    const size_t v0 = (size_t)blockIdx.x * seg;
    const size_t v1 = v0 + seg < rv.nvec ? v0 + seg : rv.nvec;
    const float lse = LOG ? logf(S) : 0.f;
    const float inv = LOG ? 0.f : __frcp_rn(S);
    const uint64_t drop = l2_policy_drop();
    for (size_t base = v0; base < v1; base += (size_t)kThreads * U) {
        float4 x[U];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const size_t v = base + (size_t)j * kThreads + threadIdx.x;
            x[j] = v < v1 ? ld_stream_hint(rv.vsrc + v, drop) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const size_t v = base + (size_t)j * kThreads + threadIdx.x;
            if (v < v1) st_stream(rv.vdst + v, normalise_raw4<LOG>(x[j], M, lse, inv));
        }
    }
    if (WIN && blockIdx.x == 0) {
        const float xe = rv.edge_load(threadIdx.x, -INFINITY);
        rv.edge_store(threadIdx.x, LOG ? (xe - M) - lse : expf(xe - M) * inv);
    }
}

template <bool LOG, bool WIN>
static int launch_split(const float* a, float* out, size_t rows, size_t cols, unsigned mis0, int sm_count, cudaStream_t s) {
    const size_t nvec = cols / 4;                             // upper bound of a row's body
    const size_t unit = (size_t)kThreads * 4;                 // vectors one CTA consumes per loop step
    size_t P = (nvec + 2 * unit - 1) / (2 * unit);            // at least two steps per segment
    const size_t cap = (size_t)sm_count * 8 / rows;           // ~8 resident CTAs per SM over all rows
    if (P > cap) P = cap;
    if (P < 1) P = 1;
    const size_t seg = ((nvec + P - 1) / P + unit - 1) / unit * unit;   // whole loop steps -> aligned segment starts
    P = (nvec + seg - 1) / seg;
    if (P < 1) P = 1;
    float2* ws = nullptr;
    TRN_TRY(scratch_alloc(reinterpret_cast<void**>(&ws), rows * P * sizeof(float2), s));
    const dim3 grid((unsigned)P, (unsigned)rows);
    softmax_split_stats_kernel<WIN><<<grid, kThreads, 0, s>>>(a, rows, cols, mis0, seg, ws);
    softmax_split_write_kernel<LOG, WIN><<<grid, kThreads, 0, s>>>(a, out, rows, cols, mis0, seg, ws, (unsigned)P);
    count_launch(2);
    const cudaError_t e = cudaGetLastError();
    TRN_TRY(scratch_free(ws, s));
    TRN_CUDA(e);
    return TRN_OK;
}

// ---- one vector sharded over several GPUs ("a single 2^30 vector softmax would need two exchanges") ----
// The split-row kernels with the exchange in the middle: every rank folds its slice to ONE (max, sum) pair
// (launch_softmax_slice_stats), the ranks all_gather their pairs (8 bytes each), and every rank folds the gathered
// pairs in rank order inside its write kernel (launch_softmax_slice_apply) — so all ranks normalise by the same bits.
__global__ void __launch_bounds__(kThreads)
softmax_fold_pairs_kernel(const float2* __restrict__ ws, unsigned P, float2* __restrict__ out) {
    __shared__ float s_w[kThreads / 32];
    float m = -INFINITY, s = 0.f;
    for (unsigned p = threadIdx.x; p < P; p += kThreads) {
        const float2 q = ws[p];
        online_merge(m, s, q.x, q.y);
    }
    const float M = block_max_256(m, s_w);
    const float ref = M == -INFINITY ? 0.f : M;
    const float S = block_sum_256(s * expf(m - ref), s_w);
    if (threadIdx.x == 0) *out = make_float2(M, S);
}

static void split_geometry(size_t n, int sm_count, size_t& P, size_t& seg) {
    const size_t nvec = n / 4;
    const size_t unit = (size_t)kThreads * 4;
    P = (nvec + 2 * unit - 1) / (2 * unit);
    const size_t cap = (size_t)sm_count * 8;
    if (P > cap) P = cap;
    if (P < 1) P = 1;
    seg = ((nvec + P - 1) / P + unit - 1) / unit * unit;
    if (seg == 0) seg = unit;   // a slice shorter than one vector: edge elements only
    P = (nvec + seg - 1) / seg;
    if (P < 1) P = 1;
}

int launch_softmax_slice_stats(const float* a, size_t n, float* pair_out, cudaStream_t s) {
    Context* c = ctx();
    if (!c) return TRN_GPU_ERROR;
    if (reinterpret_cast<uintptr_t>(a) & 3u) return fail(TRN_INVALID_INPUT, "slice pointer is not 4-byte aligned");
    const unsigned mis = (unsigned)((reinterpret_cast<uintptr_t>(a) >> 2) & 3u);
    size_t P, seg;
    split_geometry(n, c->sm_count, P, seg);
    float2* ws = nullptr;
    TRN_TRY(scratch_alloc(reinterpret_cast<void**>(&ws), P * sizeof(float2), s));
    const dim3 grid((unsigned)P, 1);
    if (mis == 0 && n % 4 == 0) softmax_split_stats_kernel<false><<<grid, kThreads, 0, s>>>(a, 1, n, 0u, seg, ws);
    else                        softmax_split_stats_kernel<true><<<grid, kThreads, 0, s>>>(a, 1, n, mis, seg, ws);
    softmax_fold_pairs_kernel<<<1, kThreads, 0, s>>>(ws, (unsigned)P, reinterpret_cast<float2*>(pair_out));
    count_launch(2);
    const cudaError_t e = cudaGetLastError();
    TRN_TRY(scratch_free(ws, s));
    TRN_CUDA(e);
    return TRN_OK;
}

int launch_softmax_slice_apply(const float* a, size_t n, const float* pairs, size_t npairs, int log_variant, float* out,
                               cudaStream_t s) {
    Context* c = ctx();
    if (!c) return TRN_GPU_ERROR;
    if ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out)) & 3u)
        return fail(TRN_INVALID_INPUT, "slice pointer is not 4-byte aligned");
    const unsigned mis = (unsigned)((reinterpret_cast<uintptr_t>(a) >> 2) & 3u);
    if (mis != (unsigned)((reinterpret_cast<uintptr_t>(out) >> 2) & 3u))
        return fail(TRN_INVALID_INPUT, "input and output slices must share their alignment modulo 16 bytes");
    if (reinterpret_cast<uintptr_t>(pairs) & 7u) return fail(TRN_INVALID_INPUT, "pairs pointer is not 8-byte aligned");
    size_t P, seg;
    split_geometry(n, c->sm_count, P, seg);
    const dim3 grid((unsigned)P, 1);
    const float2* ws = reinterpret_cast<const float2*>(pairs);
    const bool win = !(mis == 0 && n % 4 == 0);
    if (log_variant) {
        if (win) softmax_split_write_kernel<true, true><<<grid, kThreads, 0, s>>>(a, out, 1, n, mis, seg, ws, (unsigned)npairs);
        else     softmax_split_write_kernel<true, false><<<grid, kThreads, 0, s>>>(a, out, 1, n, 0u, seg, ws, (unsigned)npairs);
    } else {
        if (win) softmax_split_write_kernel<false, true><<<grid, kThreads, 0, s>>>(a, out, 1, n, mis, seg, ws, (unsigned)npairs);
        else     softmax_split_write_kernel<false, false><<<grid, kThreads, 0, s>>>(a, out, 1, n, 0u, seg, ws, (unsigned)npairs);
    }
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

// Fallback: input and output misaligned differently.  One CTA per row, three passes (max, exp-sum, write);
// passes 2 and 3 hit L2 for rows that fit there.
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

template <typename K>
static int launch_clustered(K kernel, int cs, const float* a, float* out, size_t rows, size_t cols, unsigned mis0,
                            cudaStream_t s) {
    // flat grid, one row per cluster — measured faster than a capped persistent grid for the same reason as the
    // map kernels; capping the CTAs resident per SM (fewer rows in flight against L2) measured slower at every length
    size_t want = (size_t)0x7FFFFFFF / cs;
    size_t clusters = rows < want ? rows : want;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * cs));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TRN_CUDA(cudaLaunchKernelEx(&cfg, kernel, a, out, rows, cols, mis0));
    count_launch();
    return TRN_OK;
}

static bool force_generic() {
    static const bool v = [] { const char* e = getenv("TRN_ROWS_GENERIC"); return e && e[0] == '1'; }();
    return v;
}
static int env_int(const char* name) {   // tuning knob for scripts/exp/exp_long_rows.py; 0 = unset
    const char* e = getenv(name);
    return e ? atoi(e) : 0;
}
constexpr int kRing2Default = 0;   // the CTA-pair ring kernel is opt-in until it has beaten the single-CTA ring on the part

template <bool LOG, bool WIN>
static int launch_long(int cs, const float* a, float* out, size_t rows, size_t cols, unsigned mis0, cudaStream_t s) {
    // L2 eviction hints (pass 1 evict_last, pass 2 evict_first) while the rows in flight can stay in L2: 65 536 columns
    // 4.8 -> 5.9 TB/s, 128 256: 4.8 -> 5.4, 262 144: 4.5 -> 4.9; from ~3 MB rows on they thrash (2^20 columns 4.04 -> 3.94)
    const bool hint = env_int("TRN_ROWS_L2HINT") != 1 && cols < 786432;   // knob for scripts/exp/exp_long_rows.py: 1 = plain loads
#define TRN_LONG(CS_)                                                                                                   \
    return hint ? launch_clustered(softmax_rows_long_kernel<CS_, LOG, WIN, true>, CS_, a, out, rows, cols, mis0, s)     \
                : launch_clustered(softmax_rows_long_kernel<CS_, LOG, WIN, false>, CS_, a, out, rows, cols, mis0, s)
    switch (cs) {
        case 1: TRN_LONG(1);
        case 2: TRN_LONG(2);
        case 4: TRN_LONG(4);
        default: TRN_LONG(8);
    }
#undef TRN_LONG
}

template <bool LOG, bool WIN>
static int dispatch_vec(const float* a, float* out, size_t rows, size_t cols, unsigned mis0, int sm_count, cudaStream_t s) {
    const size_t nvec = cols / 4;   // body vectors per row (the body of a window row is never longer)
    // smallest configuration whose register slots hold the row (see the file header)
    if (nvec <= kThreads * 1)      return launch_rows_warp<LOG ? 1 : 0, WIN>(a, out, rows, cols, nullptr, nullptr, 0.f, mis0, s);
    if (nvec <= kThreads * 2)      return launch_cta<256, 2, LOG, WIN>(a, out, rows, cols, mis0, s);
    if (nvec <= kThreads * 4)      return launch_cta<256, 4, LOG, WIN>(a, out, rows, cols, mis0, s);
    if (nvec <= kThreads * 8)      return launch_cta<256, 8, LOG, WIN>(a, out, rows, cols, mis0, s);
    if (nvec <= 512 * 8)           return launch_cta<512, 8, LOG, WIN>(a, out, rows, cols, mis0, s);
    // (a 1024-thread CTA holding a 32 000-float row measured 4.9 TB/s against the ring kernel's 6.0: one CTA
    //  per SM leaves nothing to overlap a row's load phase with)
    // Rows longer than the register kernels hold (scripts/exp/exp_ring.py, scripts/exp/exp_long_cs.py):
    //   * aligned rows of 28 672 < cols <= 32 768 (config 5: 32 000): the TMA ring kernel (5.98 / 6.13 TB/s at
    //     4096 x 32 000, 6.12 / 6.43 at 8192 rows; the two-pass kernel ties at 4096 rows and is 2 % behind at 8192);
    //   * a few long rows: the split-row kernels;
    //   * everything else: the two-pass cluster kernel, cluster size by row length — 1 CTA up to 20 480 columns
    //     (17 000: 6.0 TB/s where the ring reached 4.1-4.4), 2 up to 43 008 (24 576: 6.15 vs 5.2-5.6), 4 up to 100 000,
    //     8 beyond, 4 again from 786 432 (rows that no longer fit L2 between the passes).
    const int force_cs = env_int("TRN_ROWS_LONG_CS");
    const int force_hpc = env_int("TRN_RING_HPC");   // experiment knobs, read per call
    // TRN_RING2 (experiment knob, read per call): 1 forces the CTA-pair ring for aligned rows of 28 672 < cols <= 32 768
    // with cols % 8 == 0, 0 keeps the single-CTA ring
    const int ring2_mode = env_int_default("TRN_RING2", kRing2Default);
    if (!force_cs && !force_hpc && !WIN && ring2_mode && cols > 28672 && cols <= 32768 && cols % 8 == 0)
        return launch_ring2<LOG>(a, out, rows, cols, sm_count, s);
    if (!force_cs && (force_hpc || (!WIN && cols > 28672)) && nvec <= (size_t)ring::kRowVec) {
        // softmax: 32 KiB slots (7 of them) beat 16 KiB ones by 2-5 % in every run (5.98-6.15 vs 5.67-6.06 TB/s at
        // 32 000 columns).  log_softmax is bimodal with 32 KiB slots from one box visit to the next (5.56-5.68 or
        // 6.13-6.43) and steady with 16 KiB ones (5.87-6.08): it keeps 16 KiB.
        const int hpc = force_hpc ? force_hpc : LOG ? 2 : 4;
        switch (hpc) {
            case 1: return launch_ring<LOG, WIN, 1>(a, out, rows, cols, mis0, sm_count, s);
            case 4: return launch_ring<LOG, WIN, 4>(a, out, rows, cols, mis0, sm_count, s);
            default: return launch_ring<LOG, WIN, 2>(a, out, rows, cols, mis0, sm_count, s);
        }
    }
    const int cs = force_cs ? force_cs : cols <= 20480 ? 1 : cols <= 43008 ? 2 : (cols <= 100000 || cols >= 786432) ? 4 : 8;
    if (!force_cs && rows * (size_t)cs < 2 * (size_t)sm_count && cols > 32768)   // fewer than two CTAs per SM: split the rows
        return launch_split<LOG, WIN>(a, out, rows, cols, mis0, sm_count, s);
    return launch_long<LOG, WIN>(cs, a, out, rows, cols, mis0, s);
}

template <bool LOG>
static int dispatch(const float* a, float* out, size_t rows, size_t cols, int sm_count, cudaStream_t s) {
    // misalignment of the first row in elements; rows are addressable as an aligned body plus edge elements when
    // input and output share it
    const unsigned mis_in = (unsigned)((reinterpret_cast<uintptr_t>(a) >> 2) & 3u);
    const unsigned mis_out = (unsigned)((reinterpret_cast<uintptr_t>(out) >> 2) & 3u);
    const bool word_ok = ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out)) & 3u) == 0;
    if (word_ok && mis_in == mis_out && !force_generic()) {
        if (cols % 4 == 0 && mis_in == 0) return dispatch_vec<LOG, false>(a, out, rows, cols, 0u, sm_count, s);
        return dispatch_vec<LOG, true>(a, out, rows, cols, mis_in, sm_count, s);
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
            dst[i] = gamma ? __fadd_rn(__fmul_rn(__fmul_rn(gamma[i], __fsub_rn(src[i], mean)), inv_std), beta[i])
                           : __fmul_rn(__fsub_rn(src[i], mean), inv_std);   // layer_norm_simple (src/vector.rs:1386)
    }
}

// layer_norm for rows longer than the register kernels hold (cols > 16 384, 16-byte aligned): a row over a CS-CTA
// cluster, THREE streaming passes with the reference's own two-pass variance — sum -> mean, sum((x - mean)^2) -> variance,
// then the per-element formula — of which only the first reads HBM: it parks the row in L2 (evict_last), passes 2 and 3
// re-read it from there (the last one with evict_first).  Block trees + a rank-order fold through distributed shared
// memory -> deterministic.  Replaces the one-CTA-per-row scalar fallback for these rows (2.7-3.6 TB/s).
template <int CS>
__global__ void __launch_bounds__(kThreads)
layer_norm_rows_long_kernel(const float* __restrict__ in, const float* __restrict__ gamma, const float* __restrict__ beta,
                            float eps, float* __restrict__ out, size_t rows, size_t cols) {
    constexpr int U = 4;
    __shared__ float s_w[kThreads / 32];
    __shared__ float s_stat[2];
    unsigned rank = 0;
    if (CS > 1) rank = cg::this_cluster().block_rank();
    const size_t cluster_id = blockIdx.x / CS, num_clusters = gridDim.x / CS;
    const uint64_t keep = l2_policy_keep(), drop = l2_policy_drop();
    const size_t nvec = cols >> 2;
    const size_t step = (size_t)CS * kThreads * U;
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);

    // sum over the cluster of a per-CTA value, folded in rank order by every CTA; slot 0 / 1 alternate between the passes
    auto cluster_sum = [&](float v, int slot) {
        float r = block_sum_256(v, s_w);
        if (CS > 1) {
            cg::cluster_group cluster = cg::this_cluster();
            if (threadIdx.x == 0) s_stat[slot] = r;
            cluster.sync();
            float t = 0.f;
#pragma unroll
            for (int p = 0; p < CS; ++p) t += *cluster.map_shared_rank(&s_stat[slot], p);
            r = t;
        }
        return r;
    };

    for (size_t row = cluster_id; row < rows; row += num_clusters) {
        const float4* src = reinterpret_cast<const float4*>(in + row * cols);
        float4* dst = reinterpret_cast<float4*>(out + row * cols);
        // ---- pass 1 (HBM): sum -> mean
        float part = 0.f;
        for (size_t base = 0; base < nvec; base += step) {
            float4 x[U];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const size_t v = base + ((size_t)j * CS + rank) * kThreads + threadIdx.x;
                x[j] = v < nvec ? ld_stream_hint(src + v, keep) : zero4;
            }
#pragma unroll
            for (int j = 0; j < U; ++j) part += (x[j].x + x[j].y) + (x[j].z + x[j].w);
        }
        const float mean = cluster_sum(part, 0) / (float)cols;
        // ---- pass 2 (L2): sum((x - mean)^2) -> variance
        part = 0.f;
        for (size_t base = 0; base < nvec; base += step) {
            float4 x[U];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const size_t v = base + ((size_t)j * CS + rank) * kThreads + threadIdx.x;
                x[j] = v < nvec ? ld_stream_hint(src + v, keep) : make_float4(mean, mean, mean, mean);
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const float a = __fsub_rn(x[j].x, mean), b = __fsub_rn(x[j].y, mean);
                const float c = __fsub_rn(x[j].z, mean), d = __fsub_rn(x[j].w, mean);
                part += (__fmul_rn(a, a) + __fmul_rn(b, b)) + (__fmul_rn(c, c) + __fmul_rn(d, d));
            }
        }
        const float inv_std = 1.0f / sqrtf(cluster_sum(part, 1) / (float)cols + eps);
        // ---- pass 3 (L2): gamma * (x - mean) * inv_std + beta, evaluated in that order, unfused
        for (size_t base = 0; base < nvec; base += step) {
            float4 x[U];
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const size_t v = base + ((size_t)j * CS + rank) * kThreads + threadIdx.x;
                x[j] = v < nvec ? ld_stream_hint(src + v, drop) : zero4;
            }
#pragma unroll
            for (int j = 0; j < U; ++j) {
                const size_t v = base + ((size_t)j * CS + rank) * kThreads + threadIdx.x;
                if (v < nvec) {
                    float4 y;
                    if (gamma) {
                        const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + v);
                        const float4 bt = __ldg(reinterpret_cast<const float4*>(beta) + v);
                        y.x = __fadd_rn(__fmul_rn(__fmul_rn(g.x, __fsub_rn(x[j].x, mean)), inv_std), bt.x);
                        y.y = __fadd_rn(__fmul_rn(__fmul_rn(g.y, __fsub_rn(x[j].y, mean)), inv_std), bt.y);
                        y.z = __fadd_rn(__fmul_rn(__fmul_rn(g.z, __fsub_rn(x[j].z, mean)), inv_std), bt.z);
                        y.w = __fadd_rn(__fmul_rn(__fmul_rn(g.w, __fsub_rn(x[j].w, mean)), inv_std), bt.w);
                    } else {
                        y.x = __fmul_rn(__fsub_rn(x[j].x, mean), inv_std);
                        y.y = __fmul_rn(__fsub_rn(x[j].y, mean), inv_std);
                        y.z = __fmul_rn(__fsub_rn(x[j].z, mean), inv_std);
                        y.w = __fmul_rn(__fsub_rn(x[j].w, mean), inv_std);
                    }
                    st_stream(dst + v, y);
                }
            }
        }
        // peers must have read both s_stat slots of this row before the next row's pass 1 overwrites slot 0
        if (CS > 1) cg::this_cluster().sync();
    }
}

template <int CS>
static int launch_layer_norm_long(const float* a, const float* gamma, const float* beta, float eps, float* out, size_t rows,
                                  size_t cols, cudaStream_t s) {
    size_t want = (size_t)0x7FFFFFFF / CS;
    size_t clusters = rows < want ? rows : want;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * CS));
    cfg.blockDim = dim3(kThreads);
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TRN_CUDA(cudaLaunchKernelEx(&cfg, layer_norm_rows_long_kernel<CS>, a, gamma, beta, eps, out, rows, cols));
    count_launch();
    return TRN_OK;
}

int launch_layer_norm_rows(const float* a, const float* gamma, const float* beta, float eps, float* out, size_t rows,
                           size_t cols, cudaStream_t s) {
    Context* c = ctx();
    if (!c) return TRN_GPU_ERROR;
    if (rows == 0 || cols == 0) return TRN_OK;
    const size_t cap = (size_t)c->sm_count * 8;
    const bool vec_ok = (cols % 4 == 0) && (((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(out) |
                                              reinterpret_cast<uintptr_t>(gamma) | reinterpret_cast<uintptr_t>(beta)) & 15u) == 0);
    const size_t nvec = cols / 4;
    if (vec_ok && nvec <= 256) return launch_rows_warp<2, false>(a, out, rows, cols, gamma, beta, eps, 0u, s);
    if (vec_ok && nvec <= (size_t)512 * 8) {
        if (rows > 0x7FFFFFFFull) return fail(TRN_INVALID_INPUT, "%zu rows exceed the launch grid", rows);
        const unsigned grid = (unsigned)rows;   // flat: one row per CTA
        if (nvec <= kThreads * 2)      layer_norm_rows_reg_kernel<256, 2><<<grid, 256, 0, s>>>(a, gamma, beta, eps, out, rows, cols);
        else if (nvec <= kThreads * 4) layer_norm_rows_reg_kernel<256, 4><<<grid, 256, 0, s>>>(a, gamma, beta, eps, out, rows, cols);
        else if (nvec <= kThreads * 8) layer_norm_rows_reg_kernel<256, 8><<<grid, 256, 0, s>>>(a, gamma, beta, eps, out, rows, cols);
        else                           layer_norm_rows_reg_kernel<512, 8><<<grid, 512, 0, s>>>(a, gamma, beta, eps, out, rows, cols);
        count_launch();
        TRN_CUDA(cudaGetLastError());
        return TRN_OK;
    }
    if (vec_ok) {   // long aligned rows: three streaming passes over a cluster, two of them from L2 (cluster size as for softmax)
        if (cols <= 20480)  return launch_layer_norm_long<1>(a, gamma, beta, eps, out, rows, cols, s);
        if (cols <= 43008)  return launch_layer_norm_long<2>(a, gamma, beta, eps, out, rows, cols, s);
        if (cols <= 100000) return launch_layer_norm_long<4>(a, gamma, beta, eps, out, rows, cols, s);
        return launch_layer_norm_long<8>(a, gamma, beta, eps, out, rows, cols, s);
    }
    layer_norm_rows_kernel<<<(unsigned)(rows < cap ? rows : cap), kThreads, 0, s>>>(a, gamma, beta, eps, out, rows, cols);
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

}  // namespace trn
