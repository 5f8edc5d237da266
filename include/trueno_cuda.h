/*
 * trueno_cuda.h — C ABI of the B200 (sm_100a) backend for paiml/trueno's data-parallel hot path.
 *
 * This is the drop-in boundary: the symbols below are what trueno's `src/backends/cuda` backend
 * (new) binds over FFI.  Every entry point cites the reference interface it
 * replaces (paths relative to the trueno repository).  Plain pointers and sizes only; no C++,
 * torch or CUDA types appear in any signature (`stream` is an opaque cudaStream_t, NULL = the
 * backend's own stream).
 *
 * Conventions
 *   - Every function returns a trn_status.  0 = OK.  The non-zero values map 1:1 onto the
 *     variants of `TruenoError` (src/error.rs:8-41).  The message (byte-identical to the text the
 *     reference formats) is fetched with trn_last_error(); for TRN_SIZE_MISMATCH the two fields
 *     of `SizeMismatch { expected, actual }` come from trn_last_mismatch().  Error state is
 *     thread-local.
 *   - There is NO CPU fallback.  If no CUDA device is usable every compute call fails with
 *     TRN_GPU_ERROR (TruenoError::GpuError), it never computes on the host.
 *   - Host-slice functions (`trn_*_f32`) mirror `trait VectorBackend` (src/backends/mod.rs:52-385)
 *     and the `GpuBackend::matmul` hook (src/backends/gpu/mod.rs:434): inputs are borrowed host
 *     slices, outputs are caller-allocated host memory; the call returns when the result is in
 *     host memory.  Host memory obtained from trn_host_alloc() is pinned and is copied without
 *     an intermediate staging pass.
 *   - Device-resident functions (`trn_*_f32_dev`) take device pointers (trn_buf_ptr() or any
 *     CUDA allocation of the current device), are stream-ordered and return without
 *     synchronising, so chains of ops stay in HBM (the reference's precedent for this is
 *     `GpuCommandBatch`, src/backends/gpu/batch.rs:54-135).  Scalar results are written to
 *     device memory (`*_out` device pointers).  Reductions and the persistent row kernels keep
 *     small per-STREAM state in HBM (block partials, tickets, row-claim counters) that is valid
 *     across launches because launches on one stream are ordered: `_dev` calls captured into a
 *     CUDA graph on stream S must be replayed on S (or at least never concurrently with other
 *     trueno work on S); the first call on a stream allocates that state and therefore must not
 *     happen inside a capture (warm the stream up with one call before capturing).
 *   - All functions are thread-safe.  One process drives one GPU (trn_cuda_init(device)); the
 *     multi-GPU layer is one process per GPU (trueno_b200/parallel.py, torch.distributed/NCCL).
 *   - Matrices are dense row-major f32, exactly like `Matrix<f32>` (src/matrix.rs:49-54).
 */
#ifndef TRUENO_CUDA_H
#define TRUENO_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define TRN_API __attribute__((visibility("default")))
#else
#define TRN_API
#endif

/* TruenoError (src/error.rs:8-41) */
typedef enum trn_status {
    TRN_OK = 0,
    TRN_SIZE_MISMATCH = 1,       /* SizeMismatch { expected, actual } */
    TRN_INVALID_INPUT = 2,       /* InvalidInput(String) */
    TRN_EMPTY_VECTOR = 3,        /* EmptyVector */
    TRN_DIVISION_BY_ZERO = 4,    /* DivisionByZero (zscore / minmax_normalize / correlation on constant vectors) */
    TRN_GPU_ERROR = 5,           /* GpuError(String) — any CUDA failure; never a fallback */
    TRN_UNSUPPORTED_BACKEND = 6  /* UnsupportedBackend(Backend) — not an sm_100 device */
} trn_status;

/* ---- context ------------------------------------------------------------------------------- */
/* Replaces GpuBackend::new()/GpuDevice::new() (src/backends/gpu/mod.rs:63-120, device.rs:31):
 * binds this process to `device` (use -1 for "current / LOCAL_RANK / 0") once; later calls are
 * no-ops when the device matches.  Every compute entry point initialises lazily with -1. */
TRN_API int trn_cuda_init(int device);
TRN_API int trn_cuda_shutdown(void);
/* GpuBackend::is_available() (src/backends/gpu/mod.rs:75): 1 if an sm_100 device is usable. */
TRN_API int trn_cuda_is_available(void);
TRN_API int trn_device_count(int* count);
/* name (NUL-terminated, truncated to cap), SM count, total HBM bytes of the bound device */
TRN_API int trn_device_info(char* name, size_t cap, int* sm_count, uint64_t* hbm_bytes);
/* Copies the calling thread's last error message (NUL-terminated) and returns its full length. */
TRN_API size_t trn_last_error(char* buf, size_t cap);
TRN_API void trn_last_mismatch(uint64_t* expected, uint64_t* actual);
/* Blocks until all work queued on `stream` (NULL = backend stream) has finished. */
TRN_API int trn_synchronize(void* stream);
/* Number of kernels this library has launched since init (bench.py reports it as gpu_launches). */
TRN_API uint64_t trn_launch_count(void);

/* ---- device buffers with pinned host staging ------------------------------------------------
 * The device-buffer type of the north star.  Precedent: trueno-gpu GpuBuffer<T>
 * (trueno-gpu/src/driver/memory.rs:53-164) and GpuCommandBatch::upload/read
 * (src/backends/gpu/batch.rs:140-200). */
typedef struct trn_buf trn_buf;
TRN_API int trn_buf_alloc(size_t len, trn_buf** out);              /* len f32 elements in HBM */
TRN_API int trn_buf_free(trn_buf* buf);
TRN_API int trn_buf_upload(trn_buf* buf, const float* host, size_t len);
TRN_API int trn_buf_download(const trn_buf* buf, float* host, size_t len);
TRN_API size_t trn_buf_len(const trn_buf* buf);
TRN_API float* trn_buf_ptr(const trn_buf* buf);                    /* device pointer */
/* Pinned host memory (page-locked, 256-byte aligned) for zero-staging host-slice calls. */
TRN_API int trn_host_alloc(size_t len, float** out);
TRN_API int trn_host_free(float* ptr);

/* ---- Vector reductions: host slices ---------------------------------------------------------
 * trait VectorBackend::{dot,sum,max,min,argmax,argmin,norm_l2} (src/backends/mod.rs) behind
 * Vector::{dot,sum,max,min,argmax,argmin,norm_l2} (src/vector.rs:588,635,653,701,749,797,2601).
 * Validation and messages follow src/vector.rs:589-594,654-656,702-704,750-752,798-800,2602-2604:
 *   dot: na != nb -> TRN_SIZE_MISMATCH{expected=na, actual=nb}; empty -> 0
 *   sum: empty -> 0 ; norm_l2: empty -> 0
 *   max/min/argmax/argmin: empty -> TRN_INVALID_INPUT "Empty vector"
 * max/min/argmax/argmin implement the scalar backend's rule (src/backends/scalar.rs:112-166):
 * seed with a[0], strict compare, first occurrence, 64-bit indices. */
TRN_API int trn_dot_f32(const float* a, size_t na, const float* b, size_t nb, float* out);
TRN_API int trn_sum_f32(const float* a, size_t n, float* out);
TRN_API int trn_max_f32(const float* a, size_t n, float* out);
TRN_API int trn_min_f32(const float* a, size_t n, float* out);
TRN_API int trn_argmax_f32(const float* a, size_t n, uint64_t* out);
TRN_API int trn_argmin_f32(const float* a, size_t n, uint64_t* out);
TRN_API int trn_norm_l2_f32(const float* a, size_t n, float* out);

/* ---- Elementwise maps: host slices ----------------------------------------------------------
 * VectorBackend::{add,mul,sigmoid,gelu} behind Vector::{add,mul,sigmoid,gelu}
 * (src/vector.rs:358,478,1854,2179).  add/mul: na != nb -> TRN_SIZE_MISMATCH; bit-exact.
 * sigmoid/gelu: empty -> TRN_EMPTY_VECTOR; scalar-backend definitions
 * (src/backends/scalar.rs:313-340). */
TRN_API int trn_add_f32(const float* a, size_t na, const float* b, size_t nb, float* out);
TRN_API int trn_mul_f32(const float* a, size_t na, const float* b, size_t nb, float* out);
TRN_API int trn_sigmoid_f32(const float* a, size_t n, float* out);
TRN_API int trn_gelu_f32(const float* a, size_t n, float* out);

/* ---- softmax / log_softmax ------------------------------------------------------------------
 * Vector::softmax / log_softmax (src/vector.rs:1516,1581) and the GPU hook
 * GpuDevice::softmax/log_softmax (src/backends/gpu/device.rs:951,980).  `rows` independent
 * vectors of `cols` elements each, contiguous (rows == 1 is exactly Vector::softmax).
 * cols == 0 (or rows == 0) -> TRN_EMPTY_VECTOR. */
TRN_API int trn_softmax_rows_f32(const float* a, float* out, size_t rows, size_t cols);
TRN_API int trn_log_softmax_rows_f32(const float* a, float* out, size_t rows, size_t cols);

/* ---- Matrix products: host slices -----------------------------------------------------------
 * Matrix::matmul (src/matrix.rs:285) / GpuBackend::matmul(a, b, m, k, n)
 * (src/backends/gpu/mod.rs:434).  A is a_rows x a_cols, B is b_rows x b_cols, C is
 * a_rows x b_cols.  a_cols != b_rows -> TRN_INVALID_INPUT
 * "Matrix dimension mismatch for multiplication: {}x{} x {}x{} (inner dimensions {} and {} must match)"
 * (with U+00D7 multiplication signs, as the reference formats it, src/matrix.rs:286-291). */
TRN_API int trn_matmul_f32(const float* a, size_t a_rows, size_t a_cols,
                           const float* b, size_t b_rows, size_t b_cols, float* c);
/* Matrix::batched_matmul (src/matrix.rs:383): A [batch,m,k], B [batch,k,n] -> C [batch,m,n].
 * a_len/b_len are the slice lengths; mismatch -> TRN_INVALID_INPUT "A data size mismatch: ..." */
TRN_API int trn_batched_matmul_f32(const float* a, size_t a_len, const float* b, size_t b_len, float* c,
                                   size_t batch, size_t m, size_t k, size_t n);
/* Matrix::batched_matmul_4d (src/matrix.rs:464): A [batch,heads,m,k], B [batch,heads,k,n]. */
TRN_API int trn_batched_matmul_4d_f32(const float* a, size_t a_len, const float* b, size_t b_len, float* c,
                                      size_t batch, size_t heads, size_t m, size_t k, size_t n);
/* Matrix::matvec (src/matrix.rs:1657): y = A v.  v_len != cols -> TRN_INVALID_INPUT
 * "Vector length {} does not match matrix columns {} for matrix-vector multiplication". */
TRN_API int trn_matvec_f32(const float* a, size_t rows, size_t cols, const float* v, size_t v_len, float* y);
/* Matrix::transpose (src/matrix.rs:1590) — used by callers of matmul (eigen.rs:483-485). */
TRN_API int trn_transpose_f32(const float* a, size_t rows, size_t cols, float* out);

/* ---- device-resident twins ------------------------------------------------------------------
 * Same semantics and validation; pointers are device pointers; results stay in HBM.  Scalar
 * outputs (`out`) are device pointers to one f32 / one u64.  Stream-ordered, no host sync.
 * The argmax/argmin twins also emit the winning value (`out_value`, may be NULL) so that a
 * multi-GPU caller can combine (value, index) pairs across slices. */
TRN_API int trn_dot_f32_dev(const float* a, size_t na, const float* b, size_t nb, float* out, void* stream);
TRN_API int trn_sum_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_max_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_min_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_argmax_f32_dev(const float* a, size_t n, uint64_t* out, float* out_value, void* stream);
TRN_API int trn_argmin_f32_dev(const float* a, size_t n, uint64_t* out, float* out_value, void* stream);
/* One contiguous slice of a vector sharded across GPUs (trueno_b200/parallel.py).  first_slice != 0:
 * the slice starts at global index 0 and carries the a[0] seed rule; otherwise the slice has no seed
 * and reports "no candidate" (all elements NaN or the identity) as index UINT64_MAX. */
TRN_API int trn_argmax_slice_f32_dev(const float* a, size_t n, int first_slice, uint64_t* out, float* out_value, void* stream);
TRN_API int trn_argmin_slice_f32_dev(const float* a, size_t n, int first_slice, uint64_t* out, float* out_value, void* stream);
/* Same, packed for ONE all_gather: writes {value, GLOBAL index = slice_start + local index} (index UINT64_MAX =
 * "no candidate"); slice_start == 0 carries the a[0] seed rule.  trn_arg_combine_f32_dev folds `count` gathered
 * pairs (slice 0 first) into the whole-vector answer: NaN seed wins, else best value then lowest global index
 * (NCCL has no arg-reduce: SURVEY.md 8e).  One kernel each; no host synchronisation.  An EMPTY slice with
 * slice_start > 0 (more ranks than aligned blocks) reports "no candidate" and still takes part in the exchange; an empty
 * slice 0 is an empty vector: TRN_INVALID_INPUT "Empty vector". */
typedef struct trn_arg_pair { float value; uint32_t reserved; uint64_t index; } trn_arg_pair;
TRN_API int trn_argmax_slice_pair_f32_dev(const float* a, size_t n, uint64_t slice_start, trn_arg_pair* out, void* stream);
TRN_API int trn_argmin_slice_pair_f32_dev(const float* a, size_t n, uint64_t slice_start, trn_arg_pair* out, void* stream);
TRN_API int trn_arg_combine_f32_dev(const trn_arg_pair* pairs, size_t count, int is_max, uint64_t* out_idx, float* out_value, void* stream);
TRN_API int trn_norm_l2_f32_dev(const float* a, size_t n, float* out, void* stream);
/* sum of squares without the sqrt: the per-slice partial of a sharded norm_l2 (allreduce, then sqrt) */
TRN_API int trn_sumsq_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_add_f32_dev(const float* a, size_t na, const float* b, size_t nb, float* out, void* stream);
TRN_API int trn_mul_f32_dev(const float* a, size_t na, const float* b, size_t nb, float* out, void* stream);
TRN_API int trn_sigmoid_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_gelu_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_softmax_rows_f32_dev(const float* a, float* out, size_t rows, size_t cols, void* stream);
TRN_API int trn_log_softmax_rows_f32_dev(const float* a, float* out, size_t rows, size_t cols, void* stream);
TRN_API int trn_matmul_f32_dev(const float* a, size_t a_rows, size_t a_cols,
                               const float* b, size_t b_rows, size_t b_cols, float* c, void* stream);
TRN_API int trn_batched_matmul_f32_dev(const float* a, size_t a_len, const float* b, size_t b_len, float* c,
                                       size_t batch, size_t m, size_t k, size_t n, void* stream);
TRN_API int trn_batched_matmul_4d_f32_dev(const float* a, size_t a_len, const float* b, size_t b_len, float* c,
                                          size_t batch, size_t heads, size_t m, size_t k, size_t n, void* stream);
TRN_API int trn_matvec_f32_dev(const float* a, size_t rows, size_t cols, const float* v, size_t v_len,
                               float* y, void* stream);
TRN_API int trn_transpose_f32_dev(const float* a, size_t rows, size_t cols, float* out, void* stream);

/* ---- remaining VectorBackend surface ----------------------------------
 * trait VectorBackend::{sub,div,scale,abs,clamp,lerp,fma,relu,exp,swish,tanh,sqrt,recip,ln,log2,log10,
 * sin,cos,tan,floor,ceil,round,sum_kahan,norm_l1,norm_linf} (src/backends/mod.rs:67-385) behind the Vector
 * methods of the same names (src/vector.rs:423-4182); scalar-backend definitions (src/backends/scalar.rs).
 * Host-slice form `trn_<op>_f32(...)`, device-resident twin `trn_<op>_f32_dev(..., stream)`.
 *   sub, div, lerp, fma : operand lengths differ -> TRN_SIZE_MISMATCH{expected = len(a), actual}
 *   relu, swish, tanh   : empty input -> TRN_EMPTY_VECTOR (src/vector.rs:1670, :2293, :3950)
 *   other maps          : empty input -> OK, empty result
 *   clamp               : min > max -> TRN_INVALID_INPUT "Invalid clamp range: min ({}) > max ({})"
 *   sum_kahan, norm_l1, norm_linf : empty -> 0
 *   mean, variance, stddev : empty -> TRN_EMPTY_VECTOR; variance = E[x^2] - mean^2 (src/vector.rs:973-990)
 * Bit-exact vs the scalar backend: sub, div, scale, abs, clamp, lerp, fma (unfused), relu, sqrt, recip,
 * floor, ceil, round.  Transcendentals: <= 4 ulp vs libm (swish: 6 ulp, tan: 8 ulp). */
TRN_API int trn_abs_f32(const float* a, size_t n, float* out);
TRN_API int trn_abs_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_relu_f32(const float* a, size_t n, float* out);
TRN_API int trn_relu_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_exp_f32(const float* a, size_t n, float* out);
TRN_API int trn_exp_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_swish_f32(const float* a, size_t n, float* out);
TRN_API int trn_swish_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_tanh_f32(const float* a, size_t n, float* out);
TRN_API int trn_tanh_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_sqrt_f32(const float* a, size_t n, float* out);
TRN_API int trn_sqrt_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_recip_f32(const float* a, size_t n, float* out);
TRN_API int trn_recip_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_ln_f32(const float* a, size_t n, float* out);
TRN_API int trn_ln_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_log2_f32(const float* a, size_t n, float* out);
TRN_API int trn_log2_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_log10_f32(const float* a, size_t n, float* out);
TRN_API int trn_log10_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_sin_f32(const float* a, size_t n, float* out);
TRN_API int trn_sin_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_cos_f32(const float* a, size_t n, float* out);
TRN_API int trn_cos_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_tan_f32(const float* a, size_t n, float* out);
TRN_API int trn_tan_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_floor_f32(const float* a, size_t n, float* out);
TRN_API int trn_floor_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_ceil_f32(const float* a, size_t n, float* out);
TRN_API int trn_ceil_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_round_f32(const float* a, size_t n, float* out);
TRN_API int trn_round_f32_dev(const float* a, size_t n, float* out, void* stream);
/* reductions: `out` is one f32 (host pointer / device pointer for the _dev twin) */
TRN_API int trn_sum_kahan_f32(const float* a, size_t n, float* out);
TRN_API int trn_sum_kahan_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_norm_l1_f32(const float* a, size_t n, float* out);
TRN_API int trn_norm_l1_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_norm_linf_f32(const float* a, size_t n, float* out);
TRN_API int trn_norm_linf_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_sub_f32(const float* a, size_t na, const float* b, size_t nb, float* out);
TRN_API int trn_sub_f32_dev(const float* a, size_t na, const float* b, size_t nb, float* out, void* stream);
TRN_API int trn_div_f32(const float* a, size_t na, const float* b, size_t nb, float* out);
TRN_API int trn_div_f32_dev(const float* a, size_t na, const float* b, size_t nb, float* out, void* stream);
TRN_API int trn_scale_f32(const float* a, size_t n, float scalar, float* out);
TRN_API int trn_scale_f32_dev(const float* a, size_t n, float scalar, float* out, void* stream);
TRN_API int trn_clamp_f32(const float* a, size_t n, float min_val, float max_val, float* out);
TRN_API int trn_clamp_f32_dev(const float* a, size_t n, float min_val, float max_val, float* out, void* stream);
TRN_API int trn_lerp_f32(const float* a, size_t na, const float* b, size_t nb, float t, float* out);
TRN_API int trn_lerp_f32_dev(const float* a, size_t na, const float* b, size_t nb, float t, float* out, void* stream);
TRN_API int trn_fma_f32(const float* a, size_t na, const float* b, size_t nb, const float* c, size_t nc, float* out);
TRN_API int trn_fma_f32_dev(const float* a, size_t na, const float* b, size_t nb, const float* c, size_t nc, float* out,
                            void* stream);
TRN_API int trn_mean_f32(const float* a, size_t n, float* out);
TRN_API int trn_variance_f32(const float* a, size_t n, float* out);
TRN_API int trn_stddev_f32(const float* a, size_t n, float* out);

/* ---- the rest of Vector's element-wise / statistics API (widening past SURVEY.md 8f) ----------
 * Vector methods the reference implements as scalar closures over the host buffer (src/vector.rs), some with GpuDevice
 * hooks (leaky_relu / elu / clip: src/backends/gpu/device.rs).  Same conventions as above.  Error contract per op as in
 * src/vector.rs: hardswish / mish / selu / leaky_relu / elu on an empty vector -> TRN_EMPTY_VECTOR; leaky_relu with a
 * slope outside [0, 1) -> TRN_INVALID_INPUT("negative_slope must be in [0.0, 1.0), got {}"); elu with alpha <= 0 ->
 * TRN_INVALID_INPUT("alpha must be > 0, got {}"); clip with min > max -> TRN_INVALID_INPUT("min_val ({}) must be <=
 * max_val ({})"); minimum / maximum / copysign / covariance / correlation -> TRN_SIZE_MISMATCH; zscore /
 * minmax_normalize / correlation on a constant vector -> TRN_DIVISION_BY_ZERO.  neg, signum, trunc, fract, hardswish,
 * leaky_relu, minimum, maximum, copysign, clip and minmax_normalize are bit-exact against the reference. */
TRN_API int trn_neg_f32(const float* a, size_t n, float* out);   /* Vector::neg, src/vector.rs:4399 */
TRN_API int trn_neg_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_signum_f32(const float* a, size_t n, float* out);   /* Vector::signum, src/vector.rs:4261 */
TRN_API int trn_signum_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_trunc_f32(const float* a, size_t n, float* out);   /* Vector::trunc, src/vector.rs:4211 */
TRN_API int trn_trunc_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_fract_f32(const float* a, size_t n, float* out);   /* Vector::fract, src/vector.rs:4237 */
TRN_API int trn_fract_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_sinh_f32(const float* a, size_t n, float* out);   /* Vector::sinh, src/vector.rs:3885 */
TRN_API int trn_sinh_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_cosh_f32(const float* a, size_t n, float* out);   /* Vector::cosh, src/vector.rs:3916 */
TRN_API int trn_cosh_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_asin_f32(const float* a, size_t n, float* out);   /* Vector::asin, src/vector.rs:3761 */
TRN_API int trn_asin_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_acos_f32(const float* a, size_t n, float* out);   /* Vector::acos, src/vector.rs:3807 */
TRN_API int trn_acos_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_atan_f32(const float* a, size_t n, float* out);   /* Vector::atan, src/vector.rs:3855 */
TRN_API int trn_atan_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_asinh_f32(const float* a, size_t n, float* out);   /* Vector::asinh, src/vector.rs:4059 */
TRN_API int trn_asinh_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_acosh_f32(const float* a, size_t n, float* out);   /* Vector::acosh, src/vector.rs:4089 */
TRN_API int trn_acosh_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_atanh_f32(const float* a, size_t n, float* out);   /* Vector::atanh, src/vector.rs:4111 */
TRN_API int trn_atanh_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_hardswish_f32(const float* a, size_t n, float* out);   /* Vector::hardswish, src/vector.rs:2409 */
TRN_API int trn_hardswish_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_mish_f32(const float* a, size_t n, float* out);   /* Vector::mish, src/vector.rs:2477 */
TRN_API int trn_mish_f32_dev(const float* a, size_t n, float* out, void* stream);
TRN_API int trn_selu_f32(const float* a, size_t n, float* out);   /* Vector::selu, src/vector.rs:2546 */
TRN_API int trn_selu_f32_dev(const float* a, size_t n, float* out, void* stream);
/* Vector::leaky_relu (src/vector.rs:1980; GpuDevice::leaky_relu): x > 0 ? x : negative_slope * x */
TRN_API int trn_leaky_relu_f32(const float* a, size_t n, float negative_slope, float* out);
TRN_API int trn_leaky_relu_f32_dev(const float* a, size_t n, float negative_slope, float* out, void* stream);
/* Vector::elu (src/vector.rs:2085; GpuDevice::elu): x > 0 ? x : alpha * (exp(x) - 1) */
TRN_API int trn_elu_f32(const float* a, size_t n, float alpha, float* out);
TRN_API int trn_elu_f32_dev(const float* a, size_t n, float alpha, float* out, void* stream);
/* Vector::pow (src/vector.rs:3342): powf(x, exponent); exponents 2, 0.5, -1, 0 and 1 are evaluated as the single
 * correctly rounded operation a correctly rounded powf returns (x*x, sqrt, 1/x, 1, x) */
TRN_API int trn_pow_f32(const float* a, size_t n, float exponent, float* out);
TRN_API int trn_pow_f32_dev(const float* a, size_t n, float exponent, float* out, void* stream);
/* Vector::clip (src/vector.rs:1448; GpuDevice::clip): x.max(min_val).min(max_val) */
TRN_API int trn_clip_f32(const float* a, size_t n, float min_val, float max_val, float* out);
TRN_API int trn_clip_f32_dev(const float* a, size_t n, float min_val, float max_val, float* out, void* stream);
/* Vector::minimum / maximum / copysign (src/vector.rs:4328 / :4364 / :4292) */
TRN_API int trn_minimum_f32(const float* a, size_t na, const float* b, size_t nb, float* out);
TRN_API int trn_minimum_f32_dev(const float* a, size_t na, const float* b, size_t nb, float* out, void* stream);
TRN_API int trn_maximum_f32(const float* a, size_t na, const float* b, size_t nb, float* out);
TRN_API int trn_maximum_f32_dev(const float* a, size_t na, const float* b, size_t nb, float* out, void* stream);
TRN_API int trn_copysign_f32(const float* a, size_t na, const float* b, size_t nb, float* out);
TRN_API int trn_copysign_f32_dev(const float* a, size_t na, const float* b, size_t nb, float* out, void* stream);
/* out = (a - shift) * scale: the map half of zscore / minmax_normalize for device-resident callers */
TRN_API int trn_affine_f32_dev(const float* a, size_t n, float shift, float scale, float* out, void* stream);
/* Vector::sum_of_squares (src/vector.rs:898), covariance (:1063), correlation (:1119), zscore (:1180),
 * minmax_normalize (:1248): one upload, reductions and map on the resident copy, scalars through the host as in
 * the reference's own composition */
TRN_API int trn_sum_of_squares_f32(const float* a, size_t n, float* out);
TRN_API int trn_covariance_f32(const float* a, size_t na, const float* b, size_t nb, float* out);
TRN_API int trn_correlation_f32(const float* a, size_t na, const float* b, size_t nb, float* out);
TRN_API int trn_zscore_f32(const float* a, size_t n, float* out);
TRN_API int trn_minmax_normalize_f32(const float* a, size_t n, float* out);
/* Vector::layer_norm_simple (src/vector.rs:1386): (x - mean) / sqrt(var + eps) per row, no gamma / beta */
TRN_API int trn_layer_norm_simple_rows_f32(const float* a, float eps, float* out, size_t rows, size_t cols);
TRN_API int trn_layer_norm_simple_rows_f32_dev(const float* a, float eps, float* out, size_t rows, size_t cols, void* stream);

/* Matrix::embedding_lookup (src/matrix.rs:2008-2041; embedding_lookup_sparse :2059 adds the sorted unique indices, host
 * metadata): out[r, :] = table[indices[r], :], table rows x cols row-major, out n_indices x cols.  Bit-exact (byte
 * movement).  Host-slice call: an index >= rows -> TRN_INVALID_INPUT("Index {} at position {} is out of bounds for
 * embedding table with {} rows"), checked before anything is launched; empty indices -> nothing written.  Resident
 * twin: indices are device memory (uint64), an out-of-range index yields a zero row. */
TRN_API int trn_embedding_lookup_f32(const float* table, size_t rows, size_t cols, const uint64_t* indices, size_t n_indices,
                                     float* out);
TRN_API int trn_embedding_lookup_f32_dev(const float* table, size_t rows, size_t cols, const uint64_t* indices,
                                         size_t n_indices, float* out, void* stream);

/* ---- device-resident op chaining --------------------------------------
 * The CUDA counterpart of GpuCommandBatch (src/backends/gpu/batch.rs:118-1019): queue uploads and ops, run
 * them with ONE graph launch, read results back.  Every BufferId is a slice of one device arena; execute()
 * uploads the inputs and launches the op sequence as a CUDA graph captured from the same launchers the `_dev`
 * entry points use (bit-identical results); calling execute() again (after trn_batch_update) replays it.
 * ops: 0 relu, 1 scale(scalar), 2 add, 3 mul, 4 dot (1-element result), 5 sigmoid, 6 tanh, 7 swish, 8 gelu, 9 sub.
 * Binary ops on buffers of different sizes: TRN_INVALID_INPUT "Buffer size mismatch: {} vs {}" (the reference
 * panics with the same text, batch.rs:215-232). */
typedef struct trn_batch trn_batch;
TRN_API int trn_batch_create(trn_batch** out);
TRN_API int trn_batch_destroy(trn_batch* batch);
TRN_API int trn_batch_upload(trn_batch* batch, const float* data, size_t len, uint32_t* id);
TRN_API int trn_batch_update(trn_batch* batch, uint32_t id, const float* data, size_t len);
TRN_API int trn_batch_op(trn_batch* batch, int op, uint32_t a_id, uint32_t b_id, float scalar, uint32_t* out_id);
TRN_API int trn_batch_execute(trn_batch* batch);
TRN_API int trn_batch_read(trn_batch* batch, uint32_t id, float* out, size_t len);
TRN_API size_t trn_batch_num_operations(const trn_batch* batch);
TRN_API size_t trn_batch_num_buffers(const trn_batch* batch);

/* ---- callers next to the path -----------------------------------------
 * Matrix::vecmat (src/matrix.rs:1782): y[cols] = v^T A, accumulated row by row exactly like the reference
 * (result += row_i * v[i], unfused, ascending i) -> bit-exact.  v_len != rows -> TRN_INVALID_INPUT
 * "Vector length {} does not match matrix rows {} for vector-matrix multiplication".
 * Vector::layer_norm (src/vector.rs:1316): `rows` vectors of `cols` elements sharing gamma / beta
 * (rows == 1 is the reference call): y = gamma * (x - mean) * inv_std + beta.  Empty -> TRN_EMPTY_VECTOR;
 * gamma / beta length != cols -> TRN_SIZE_MISMATCH{expected = cols, actual}. */
TRN_API int trn_vecmat_f32(const float* v, size_t v_len, const float* a, size_t rows, size_t cols, float* y);
TRN_API int trn_vecmat_f32_dev(const float* v, size_t v_len, const float* a, size_t rows, size_t cols, float* y, void* stream);
TRN_API int trn_layer_norm_rows_f32(const float* a, const float* gamma, size_t gamma_len, const float* beta, size_t beta_len,
                                    float eps, float* out, size_t rows, size_t cols);
TRN_API int trn_layer_norm_rows_f32_dev(const float* a, const float* gamma, size_t gamma_len, const float* beta,
                                        size_t beta_len, float eps, float* out, size_t rows, size_t cols, void* stream);

/* ---- ONE Vector::softmax / log_softmax sharded over several GPUs -----------------
 * (src/vector.rs:1516 / :1581 on a vector whose contiguous slices live on different GPUs.)  Step 1, per rank:
 * pair_out[0] = max of the slice, pair_out[1] = sum of exp(x - max) over the slice (device memory, 8-byte aligned).
 * The ranks all_gather their pairs (rank order).  Step 2, per rank: out = exp(x - M) / S or (x - M) - ln S with (M, S)
 * the fold of all `npairs` pairs, evaluated in rank order by every rank -> identical bits everywhere.  a and out must
 * share their alignment modulo 16 bytes.  An EMPTY slice (more ranks than aligned blocks) contributes the identity pair
 * (-inf, 0) and writes nothing, so no rank drops out of the exchange; the empty-VECTOR error (src/vector.rs:1517-1519) is
 * raised by the caller that knows the whole length. */
TRN_API int trn_softmax_slice_stats_f32_dev(const float* a, size_t n, float* pair_out, void* stream);
TRN_API int trn_softmax_slice_apply_f32_dev(const float* a, size_t n, const float* pairs, size_t npairs, int log_variant,
                                            float* out, void* stream);

/* ---- fused slice reduction + exchange over NVLink peer memory -----------------
 * One process per GPU.  Each process exports a small mailbox (trn_comm_local_handle: a 64-byte CUDA IPC
 * handle), the host all-gathers the handles once, trn_comm_create maps the peers' mailboxes.  The
 * `_allreduce` / `_allgather` reductions then need no separate collective: the last block of the slice kernel
 * writes its result into every peer's mailbox with P2P stores and folds all ranks' results in rank order, so
 * every rank gets the bit-identical whole-vector answer from ONE launch.  Collective rules: same calls, same
 * order, one stream per process; world <= 8.  The argmax/argmin variants apply the scalar-backend rule across
 * slices (NaN seed of slice 0 wins, else best value then lowest global index). */
typedef struct trn_comm trn_comm;
TRN_API int trn_comm_local_handle(void* handle64);
TRN_API int trn_comm_create(int rank, int world, const void* handles, trn_comm** out);
TRN_API int trn_comm_destroy(trn_comm* comm);
TRN_API int trn_sum_allreduce_f32_dev(trn_comm* comm, const float* a, size_t n, float* out, void* stream);
TRN_API int trn_dot_allreduce_f32_dev(trn_comm* comm, const float* a, size_t na, const float* b, size_t nb, float* out, void* stream);
TRN_API int trn_norm_l2_allreduce_f32_dev(trn_comm* comm, const float* a, size_t n, float* out, void* stream);
TRN_API int trn_argmax_allgather_f32_dev(trn_comm* comm, const float* a, size_t n, uint64_t slice_start, uint64_t* out_idx,
                                         float* out_value, void* stream);
TRN_API int trn_argmin_allgather_f32_dev(trn_comm* comm, const float* a, size_t n, uint64_t slice_start, uint64_t* out_idx,
                                         float* out_value, void* stream);
/* TRN_OK, or TRN_GPU_ERROR once an exchange on this communicator has given up waiting for a peer (TRN_PEER_TIMEOUT_MS,
 * default 30 s): a dead or desynchronised peer surfaces as TruenoError::GpuError (src/error.rs:27-28), never as a hang.
 * The timed-out call returns NaN / index UINT64_MAX; every later call on the communicator fails before launching. */
TRN_API int trn_comm_status(trn_comm* comm);

/* Matrix::convolve2d (src/matrix.rs:1868): valid-padding 2-D cross-correlation, out is
 * (rows - k_rows + 1) x (cols - k_cols + 1); accumulated in the reference's order, unfused -> bit-exact.  Kernel larger
 * than the input -> TRN_INVALID_INPUT "Kernel size ({}x{}) larger than input ({}x{})". */
TRN_API int trn_convolve2d_f32(const float* in, size_t rows, size_t cols, const float* kernel, size_t k_rows, size_t k_cols,
                               float* out);
TRN_API int trn_convolve2d_f32_dev(const float* in, size_t rows, size_t cols, const float* kernel, size_t k_rows,
                                   size_t k_cols, float* out, void* stream);

/* Fused scaled-dot-product attention: out = softmax(scale * Q K^T [causal]) V per head.
 * Replaces trueno-gpu's AttentionKernel (trueno-gpu/src/kernels/attention.rs:27-125: q/k/v/o are
 * [num_heads][seq_len][head_dim] f32, `scale` defaults to 1/sqrt(head_dim) there, `causal` masks keys after the
 * query) and the CPU composition batched_matmul_4d(Q, K^T) -> scale -> softmax -> batched_matmul_4d(P, V)
 * (src/matrix.rs:464, :3985; src/vector.rs:1516).  The score matrix is never written to memory.  head_dim <= 128
 * runs on the tensor cores (3xTF32, same accuracy contract as matmul); 128 < head_dim <= 1024 on the SIMT kernel;
 * larger -> TRN_INVALID_INPUT.  Size errors: "Q data size mismatch: expected {} (heads*seq*head_dim), got {}"
 * (likewise K, V) in the style of src/matrix.rs:481-502.  Engine follows trn_set_gemm_engine (1 = SIMT, 2 = tensor). */
TRN_API int trn_attention_f32(const float* q, size_t q_len, const float* k, size_t k_len, const float* v, size_t v_len,
                              float* out, size_t heads, size_t seq_len, size_t head_dim, float scale, int causal);
TRN_API int trn_attention_f32_dev(const float* q, size_t q_len, const float* k, size_t k_len, const float* v, size_t v_len,
                                  float* out, size_t heads, size_t seq_len, size_t head_dim, float scale, int causal,
                                  void* stream);

/* SymmetricEigen::new (src/eigen.rs:108-141; GPU hook GpuBackend::symmetric_eigen src/backends/gpu/mod.rs:466; SURVEY.md
 * 8f rank 4): eigendecomposition of a symmetric rows x cols matrix by Jacobi rotations with the reference's rotation
 * formulas, threshold (1e-7 * max(||A||_F, 1)) and sweep limit (50); the device applies the n/2 disjoint rotations of a
 * round-robin round at once.  eigenvalues[rows] come back in descending order, eigenvectors[rows*rows] row-major with
 * the eigenvectors as COLUMNS in the same order (src/eigen.rs:183-203).  Errors (TRN_INVALID_INPUT, the reference's
 * text): "Matrix must be square for eigendecomposition, got {}x{}", "Cannot compute eigendecomposition of empty
 * matrix", "Jacobi algorithm failed to converge after 50 sweeps"; rows > 8192 is rejected.  The _dev variant takes
 * and returns device pointers but synchronises the stream (the stopping rule is read back once per sweep). */
TRN_API int trn_symmetric_eigen_f32(const float* a, size_t rows, size_t cols, float* eigenvalues, float* eigenvectors);
TRN_API int trn_symmetric_eigen_f32_dev(const float* a, size_t rows, size_t cols, float* eigenvalues, float* eigenvectors,
                                        void* stream);

/* ---- pre-split right-hand operand ------------------------------------------------------------
 * Matrix::matmul(&self, other) reads `other` on every call (src/matrix.rs:285; the CPU path re-transposes it each time,
 * :934).  The tensor-core path splits B into tf32 (hi, lo) halves per call — 2 x |B| of writes; a caller that multiplies
 * many A's by ONE B (the row blocks of a sharded product src/matrix.rs:962-1011, SymmetricEigen::reconstruct
 * src/eigen.rs:483-485, weights) prepares it once.  `b` is borrowed: alive and unchanged while the handle is in use.
 * trn_matmul_prepared_f32_dev(a, handle) == trn_matmul_f32_dev(a, b) bit for bit, same errors. */
typedef struct trn_gemm_b trn_gemm_b;
TRN_API int trn_gemm_prepare_b_dev(const float* b, size_t b_rows, size_t b_cols, trn_gemm_b** out, void* stream);
TRN_API int trn_gemm_b_free(trn_gemm_b* handle);
TRN_API int trn_matmul_prepared_f32_dev(const float* a, size_t a_rows, size_t a_cols, const trn_gemm_b* b, float* c,
                                        void* stream);
/* host-slice twin: A and C are host slices (pinned ones are pipelined by row blocks), B stays resident in HBM */
TRN_API int trn_matmul_prepared_f32(const float* a, size_t a_rows, size_t a_cols, const trn_gemm_b* b, float* c);

/* ---- row blocks of ONE product (sharded Matrix::matmul) ------------------------------------------
 * src/matrix.rs:962-1011 cuts C into 256-row blocks over rayon workers; every block runs the same dot loop, so the
 * parallel product carries the bits of the sequential one.  Here `a` holds block_rows consecutive rows of an A with
 * total_rows rows (one rank's share), `c` receives the same rows of C, and the kernel is chosen as trn_matmul_f32_dev
 * chooses it for the WHOLE total_rows x a_cols x b_cols product — not by the block's own height, which would send a
 * short last block to another kernel (other rounding) and a 1-row block into the vecmat special case.  Gathered row
 * blocks == the unsharded product, bit for bit.  Errors: the reference's dimension-mismatch text on (total_rows,
 * a_cols) x (b_rows, b_cols); block_rows > total_rows is InvalidInput. */
TRN_API int trn_matmul_rowblock_f32_dev(const float* a, size_t block_rows, size_t total_rows, size_t a_cols, const float* b,
                                        size_t b_rows, size_t b_cols, float* c, void* stream);
TRN_API int trn_matmul_rowblock_prepared_f32_dev(const float* a, size_t block_rows, size_t total_rows, size_t a_cols,
                                                 const trn_gemm_b* b, float* c, void* stream);

/* ---- GEMM engine selection (measurement and tests) ------------------------------------------
 * The dispatcher picks the tcgen05 3xTF32 kernel for shapes that fill its tiles and the SIMT
 * FFMA kernel for small/skinny ones.  Tests and bench.py can force one engine to compare them
 * (BASELINE.json config 2: "3xTF32 tcgen05 vs SIMT FFMA").  0 = auto, 1 = SIMT FFMA,
 * 2 = tcgen05 3xTF32, 3 = tcgen05 1xTF32 (peak probe only; NOT fp32-accurate, never auto-selected). */
TRN_API int trn_set_gemm_engine(int engine);
TRN_API int trn_get_gemm_engine(void);

/* ---- live kernel timing (bench.py's roofline leg) -------------------------------------------
 * When enabled, the GEMM launcher brackets its operand pre-pass and its main tensor-core kernel
 * with CUDA events on the launching stream (a ring of 64 event triples: no host synchronisation
 * between calls).  trn_profile_enable(1) restarts the ring; trn_profile_last_gemm() synchronises
 * on the newest event and returns the MEAN pre-pass / kernel durations (milliseconds) over the
 * (at most 64 newest) GEMM calls recorded since.  Off by default. */
TRN_API int trn_profile_enable(int on);
TRN_API int trn_profile_last_gemm(float* prepass_ms, float* kernel_ms);

#ifdef __cplusplus
}
#endif
#endif /* TRUENO_CUDA_H */
