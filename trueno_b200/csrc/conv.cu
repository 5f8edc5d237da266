// conv.cu — Matrix::convolve2d (src/matrix.rs:1868-1950): valid-padding 2-D cross-correlation.
// SURVEY.md 8f rank 4 — one of the two other ops the reference ever sent to a GPU (src/backends/gpu/mod.rs:389).
//
// out[r][c] = sum over (kr, kc) in row-major order of in[r+kr][c+kc] * kernel[kr][kc], accumulated exactly like
// the reference's scalar loop: sum = sum + (a * b), unfused, kernel rows outer, kernel columns inner — so the
// result is BIT-EXACT against the reference for any kernel size.
//
// One 32 x 32 output tile per CTA (256 threads, 4 outputs each) on a flat grid.  The input tile plus its halo
// and the kernel sit in shared memory when they fit 47 KiB (kernels up to ~75 x 75); otherwise the same loop
// reads through L1/L2.  HBM traffic: the input once (+ halo re-reads of neighbouring tiles, from L2) and the output
// once -> 8 B per output element for small kernels; for large kernels the FMA pipe takes over (rows*cols*kr*kc
// multiply-adds).
#include "common.cuh"

namespace trn {

constexpr int kTile = 32;

template <bool SMEM>
__global__ void __launch_bounds__(256)
convolve2d_kernel(const float* __restrict__ in, const float* __restrict__ kernel, float* __restrict__ out, size_t rows,
                  size_t cols, size_t kr, size_t kc, size_t out_rows, size_t out_cols, size_t tiles_x) {
    extern __shared__ float smem[];
    const size_t ty0 = (blockIdx.x / tiles_x) * kTile, tx0 = (blockIdx.x % tiles_x) * kTile;
    const size_t tile_w = kTile + kc - 1, tile_h = kTile + kr - 1;
    float* s_in = smem;                        // [tile_h][tile_w]
    float* s_k = smem + tile_h * tile_w;       // [kr][kc]
    if (SMEM) {
        for (size_t i = threadIdx.x; i < tile_h * tile_w; i += 256) {
            const size_t r = ty0 + i / tile_w, c = tx0 + i % tile_w;
            s_in[i] = (r < rows && c < cols) ? in[r * cols + c] : 0.f;
        }
        for (size_t i = threadIdx.x; i < kr * kc; i += 256) s_k[i] = kernel[i];
        __syncthreads();
    }
    const int lx = threadIdx.x & 31, ly0 = threadIdx.x >> 5;   // column, first of 4 rows (stride 8)
    float sum[4] = {0.f, 0.f, 0.f, 0.f};
    for (size_t a = 0; a < kr; ++a) {
        for (size_t b = 0; b < kc; ++b) {
            const float kv = SMEM ? s_k[a * kc + b] : __ldg(kernel + a * kc + b);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int ly = ly0 + 8 * q;
                float v;
                if (SMEM) {
                    v = s_in[(ly + a) * tile_w + lx + b];
                } else {
                    const size_t r = ty0 + ly + a, c = tx0 + lx + b;
                    v = (r < rows && c < cols) ? __ldg(in + r * cols + c) : 0.f;
                }
                sum[q] = __fadd_rn(sum[q], __fmul_rn(v, kv));   // sum += input * kernel, unfused (src/matrix.rs:1935)
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const size_t r = ty0 + ly0 + 8 * q, c = tx0 + lx;
        if (r < out_rows && c < out_cols) out[r * out_cols + c] = sum[q];
    }
}

int launch_convolve2d(const float* in, size_t rows, size_t cols, const float* kernel, size_t kr, size_t kc, float* out,
                      cudaStream_t s) {
    Context* c = ctx();
    if (!c) return TRN_GPU_ERROR;
    const size_t out_rows = rows - kr + 1, out_cols = cols - kc + 1;
    if (out_rows == 0 || out_cols == 0 || kr == 0 || kc == 0) return TRN_OK;
    const size_t tiles_x = (out_cols + kTile - 1) / kTile, tiles_y = (out_rows + kTile - 1) / kTile;
    if (tiles_x * tiles_y > 0x7FFFFFFFull) return fail(TRN_INVALID_INPUT, "image of %zu tiles exceeds the launch grid", tiles_x * tiles_y);
    const size_t smem_bytes = ((kTile + kr - 1) * (kTile + kc - 1) + kr * kc) * sizeof(float);
    const unsigned grid = (unsigned)(tiles_x * tiles_y);
    if (smem_bytes <= 47 * 1024)
        convolve2d_kernel<true><<<grid, 256, smem_bytes, s>>>(in, kernel, out, rows, cols, kr, kc, out_rows, out_cols, tiles_x);
    else
        convolve2d_kernel<false><<<grid, 256, 0, s>>>(in, kernel, out, rows, cols, kr, kc, out_rows, out_cols, tiles_x);
    count_launch();
    TRN_CUDA(cudaGetLastError());
    return TRN_OK;
}

}  // namespace trn
