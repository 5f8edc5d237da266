//! `src/backends/cuda/mod.rs` — the B200 (sm_100a) backend of trueno's data-parallel hot path.
//!
//! Drop this directory into the trueno tree as `src/backends/cuda/`, add
//! `trueno-cuda-sys = { path = "...", optional = true }` and `cuda = ["trueno-cuda-sys"]` to
//! Cargo.toml, and apply the dispatcher edits listed in INTEGRATION.md.  With the `cuda` feature
//! on, `Backend::Auto`/`Backend::GPU` route the hot-path ops here UNCONDITIONALLY: there is no
//! size threshold, no wgpu fallback and no CPU fallback — a CUDA failure surfaces as
//! `TruenoError::GpuError`.
//!
//! The public `Vector` / `Matrix` API and its `Result` semantics do not change: shape validation
//! stays in `src/vector.rs` / `src/matrix.rs` (so the error values and message text are produced by
//! the same Rust code as today), and this backend only replaces the arithmetic.
//!
//! This file cannot be compiled in the build image of the CUDA sources (no cargo there); it is kept
//! trivially thin on purpose — every function is one FFI call plus status mapping.
use crate::{Backend, TruenoError};
use trueno_cuda_sys as sys;

pub mod buffer;
pub use buffer::DeviceBuffer;

/// Maps a non-zero `trn_status` to the `TruenoError` variant it stands for (src/error.rs:8-41).
#[cold]
pub(crate) fn status_to_error(status: i32) -> TruenoError {
    let mut buf = [0u8; 1024];
    // SAFETY: buf is writable for its whole length; the library NUL-terminates.
    let len = unsafe { sys::trn_last_error(buf.as_mut_ptr().cast(), buf.len()) }.min(buf.len() - 1);
    let msg = String::from_utf8_lossy(&buf[..len]).into_owned();
    match status {
        sys::TRN_SIZE_MISMATCH => {
            let (mut expected, mut actual) = (0u64, 0u64);
            unsafe { sys::trn_last_mismatch(&mut expected, &mut actual) };
            TruenoError::SizeMismatch { expected: expected as usize, actual: actual as usize }
        }
        sys::TRN_INVALID_INPUT => TruenoError::InvalidInput(msg),
        sys::TRN_EMPTY_VECTOR => TruenoError::EmptyVector,
        sys::TRN_DIVISION_BY_ZERO => TruenoError::DivisionByZero,
        sys::TRN_UNSUPPORTED_BACKEND => TruenoError::UnsupportedBackend(Backend::GPU),
        _ => TruenoError::GpuError(msg),
    }
}

#[inline]
pub(crate) fn check(status: i32) -> Result<(), TruenoError> {
    if status == sys::TRN_OK { Ok(()) } else { Err(status_to_error(status)) }
}

/// Stateless operator set with the shape of `trait VectorBackend` (src/backends/mod.rs:52-385),
/// but fallible: a CUDA error cannot be swallowed into a plain value.
pub struct CudaBackend;

macro_rules! reduce_f32 {
    ($name:ident, $ffi:ident) => {
        #[inline]
        pub fn $name(a: &[f32]) -> Result<f32, TruenoError> {
            let mut out = 0.0f32;
            check(unsafe { sys::$ffi(a.as_ptr(), a.len(), &mut out) })?;
            Ok(out)
        }
    };
}
macro_rules! unary_map {
    ($name:ident, $ffi:ident) => {
        #[inline]
        pub fn $name(a: &[f32], result: &mut [f32]) -> Result<(), TruenoError> {
            debug_assert_eq!(a.len(), result.len());
            check(unsafe { sys::$ffi(a.as_ptr(), a.len(), result.as_mut_ptr()) })
        }
    };
}
macro_rules! binary_map {
    ($name:ident, $ffi:ident) => {
        #[inline]
        pub fn $name(a: &[f32], b: &[f32], result: &mut [f32]) -> Result<(), TruenoError> {
            debug_assert_eq!(a.len(), result.len());
            check(unsafe { sys::$ffi(a.as_ptr(), a.len(), b.as_ptr(), b.len(), result.as_mut_ptr()) })
        }
    };
}

impl CudaBackend {
    /// `GpuBackend::is_available()` (src/backends/gpu/mod.rs:75)
    pub fn is_available() -> bool {
        unsafe { sys::trn_cuda_is_available() != 0 }
    }

    pub fn dot(a: &[f32], b: &[f32]) -> Result<f32, TruenoError> {
        let mut out = 0.0f32;
        check(unsafe { sys::trn_dot_f32(a.as_ptr(), a.len(), b.as_ptr(), b.len(), &mut out) })?;
        Ok(out)
    }
    reduce_f32!(sum, trn_sum_f32);
    reduce_f32!(max, trn_max_f32);
    reduce_f32!(min, trn_min_f32);
    reduce_f32!(norm_l2, trn_norm_l2_f32);

    pub fn argmax(a: &[f32]) -> Result<usize, TruenoError> {
        let mut out = 0u64;
        check(unsafe { sys::trn_argmax_f32(a.as_ptr(), a.len(), &mut out) })?;
        Ok(out as usize)
    }
    pub fn argmin(a: &[f32]) -> Result<usize, TruenoError> {
        let mut out = 0u64;
        check(unsafe { sys::trn_argmin_f32(a.as_ptr(), a.len(), &mut out) })?;
        Ok(out as usize)
    }

    binary_map!(add, trn_add_f32);
    binary_map!(mul, trn_mul_f32);
    unary_map!(sigmoid, trn_sigmoid_f32);
    unary_map!(gelu, trn_gelu_f32);

    // remaining `VectorBackend` maps (src/backends/mod.rs:52-385) and the rest of Vector's element-wise API
    // (src/vector.rs:1448-4410): same macro, one FFI symbol each
    unary_map!(abs, trn_abs_f32);
    unary_map!(relu, trn_relu_f32);
    unary_map!(exp, trn_exp_f32);
    unary_map!(swish, trn_swish_f32);
    unary_map!(tanh, trn_tanh_f32);
    unary_map!(sqrt, trn_sqrt_f32);
    unary_map!(recip, trn_recip_f32);
    unary_map!(ln, trn_ln_f32);
    unary_map!(log2, trn_log2_f32);
    unary_map!(log10, trn_log10_f32);
    unary_map!(sin, trn_sin_f32);
    unary_map!(cos, trn_cos_f32);
    unary_map!(tan, trn_tan_f32);
    unary_map!(floor, trn_floor_f32);
    unary_map!(ceil, trn_ceil_f32);
    unary_map!(round, trn_round_f32);
    unary_map!(neg, trn_neg_f32);
    unary_map!(signum, trn_signum_f32);
    unary_map!(trunc, trn_trunc_f32);
    unary_map!(fract, trn_fract_f32);
    unary_map!(sinh, trn_sinh_f32);
    unary_map!(cosh, trn_cosh_f32);
    unary_map!(asin, trn_asin_f32);
    unary_map!(acos, trn_acos_f32);
    unary_map!(atan, trn_atan_f32);
    unary_map!(asinh, trn_asinh_f32);
    unary_map!(acosh, trn_acosh_f32);
    unary_map!(atanh, trn_atanh_f32);
    unary_map!(hardswish, trn_hardswish_f32);
    unary_map!(mish, trn_mish_f32);
    unary_map!(selu, trn_selu_f32);
    binary_map!(sub, trn_sub_f32);
    binary_map!(div, trn_div_f32);
    binary_map!(minimum, trn_minimum_f32);
    binary_map!(maximum, trn_maximum_f32);
    binary_map!(copysign, trn_copysign_f32);
    pub fn leaky_relu(a: &[f32], result: &mut [f32], negative_slope: f32) -> Result<(), TruenoError> {
        check(unsafe { sys::trn_leaky_relu_f32(a.as_ptr(), a.len(), negative_slope, result.as_mut_ptr()) })
    }
    pub fn elu(a: &[f32], result: &mut [f32], alpha: f32) -> Result<(), TruenoError> {
        check(unsafe { sys::trn_elu_f32(a.as_ptr(), a.len(), alpha, result.as_mut_ptr()) })
    }
    pub fn pow(a: &[f32], result: &mut [f32], n: f32) -> Result<(), TruenoError> {
        check(unsafe { sys::trn_pow_f32(a.as_ptr(), a.len(), n, result.as_mut_ptr()) })
    }
    /// `GpuDevice::clip` (src/backends/gpu/device.rs) for `Vector::clip` (src/vector.rs:1448)
    pub fn clip(a: &[f32], result: &mut [f32], min_val: f32, max_val: f32) -> Result<(), TruenoError> {
        check(unsafe { sys::trn_clip_f32(a.as_ptr(), a.len(), min_val, max_val, result.as_mut_ptr()) })
    }
    unary_map!(zscore, trn_zscore_f32);
    unary_map!(minmax_normalize, trn_minmax_normalize_f32);
    reduce_f32!(sum_of_squares, trn_sum_of_squares_f32);
    pub fn covariance(a: &[f32], b: &[f32]) -> Result<f32, TruenoError> {
        let mut out = 0.0f32;
        check(unsafe { sys::trn_covariance_f32(a.as_ptr(), a.len(), b.as_ptr(), b.len(), &mut out) })?;
        Ok(out)
    }
    pub fn correlation(a: &[f32], b: &[f32]) -> Result<f32, TruenoError> {
        let mut out = 0.0f32;
        check(unsafe { sys::trn_correlation_f32(a.as_ptr(), a.len(), b.as_ptr(), b.len(), &mut out) })?;
        Ok(out)
    }
    pub fn layer_norm_simple(a: &[f32], result: &mut [f32], eps: f32) -> Result<(), TruenoError> {
        check(unsafe { sys::trn_layer_norm_simple_rows_f32(a.as_ptr(), eps, result.as_mut_ptr(), 1, a.len()) })
    }

    /// One row (`rows == 1`) is exactly `Vector::softmax` (src/vector.rs:1516).
    pub fn softmax_rows(a: &[f32], result: &mut [f32], rows: usize, cols: usize) -> Result<(), TruenoError> {
        check(unsafe { sys::trn_softmax_rows_f32(a.as_ptr(), result.as_mut_ptr(), rows, cols) })
    }
    pub fn log_softmax_rows(a: &[f32], result: &mut [f32], rows: usize, cols: usize) -> Result<(), TruenoError> {
        check(unsafe { sys::trn_log_softmax_rows_f32(a.as_ptr(), result.as_mut_ptr(), rows, cols) })
    }

    /// One `Vector::softmax` / `log_softmax` whose contiguous slices live on different GPUs (one process per GPU):
    /// step 1 folds this rank's resident slice to a (max, sum of exp) pair in device memory; the ranks all_gather their
    /// pairs (NCCL or any transport); step 2 normalises the slice by the fold of all pairs (rank order, same bits on
    /// every rank).  Raw device pointers: the slices are `DeviceBuffer`s (buffer.rs), never host memory.
    ///
    /// # Safety
    /// `slice`, `pair_out`, `pairs` and `out` must be valid device pointers on the current device.
    pub unsafe fn softmax_slice_stats(slice: *const f32, n: usize, pair_out: *mut f32, stream: *mut core::ffi::c_void)
        -> Result<(), TruenoError> {
        check(sys::trn_softmax_slice_stats_f32_dev(slice, n, pair_out, stream))
    }
    /// # Safety
    /// see `softmax_slice_stats`
    pub unsafe fn softmax_slice_apply(slice: *const f32, n: usize, pairs: *const f32, npairs: usize, log: bool, out: *mut f32,
                                      stream: *mut core::ffi::c_void) -> Result<(), TruenoError> {
        check(sys::trn_softmax_slice_apply_f32_dev(slice, n, pairs, npairs, log as i32, out, stream))
    }

    /// `GpuBackend::matmul(a, b, m, k, n)` (src/backends/gpu/mod.rs:434) — same argument meaning.
    pub fn matmul(a: &[f32], b: &[f32], m: usize, k: usize, n: usize) -> Result<Vec<f32>, TruenoError> {
        let mut c = vec![0.0f32; m * n];
        check(unsafe { sys::trn_matmul_f32(a.as_ptr(), m, k, b.as_ptr(), k, n, c.as_mut_ptr()) })?;
        Ok(c)
    }
    pub fn batched_matmul(a: &[f32], b: &[f32], batch: usize, m: usize, k: usize, n: usize) -> Result<Vec<f32>, TruenoError> {
        let mut c = vec![0.0f32; batch * m * n];
        check(unsafe {
            sys::trn_batched_matmul_f32(a.as_ptr(), a.len(), b.as_ptr(), b.len(), c.as_mut_ptr(), batch, m, k, n)
        })?;
        Ok(c)
    }
    pub fn batched_matmul_4d(a: &[f32], b: &[f32], batch: usize, heads: usize, m: usize, k: usize, n: usize)
        -> Result<Vec<f32>, TruenoError> {
        let mut c = vec![0.0f32; batch * heads * m * n];
        check(unsafe {
            sys::trn_batched_matmul_4d_f32(a.as_ptr(), a.len(), b.as_ptr(), b.len(), c.as_mut_ptr(), batch, heads, m, k, n)
        })?;
        Ok(c)
    }
    /// `GpuBackend::symmetric_eigen` (src/backends/gpu/mod.rs:466) for `SymmetricEigen::compute_gpu` (src/eigen.rs:309):
    /// returns (eigenvalues, eigenvectors as columns), already in descending order.
    pub fn symmetric_eigen(matrix: &[f32], n: usize) -> Result<(Vec<f32>, Vec<f32>), TruenoError> {
        let mut values = vec![0.0f32; n];
        let mut vectors = vec![0.0f32; n * n];
        check(unsafe { sys::trn_symmetric_eigen_f32(matrix.as_ptr(), n, n, values.as_mut_ptr(), vectors.as_mut_ptr()) })?;
        Ok((values, vectors))
    }
    /// Fused attention, the device counterpart of `trueno_gpu::kernels::AttentionKernel`
    /// (trueno-gpu/src/kernels/attention.rs:27-125): q, k, v are `[heads][seq_len][head_dim]`.
    pub fn attention(q: &[f32], k: &[f32], v: &[f32], heads: usize, seq_len: usize, head_dim: usize, scale: f32, causal: bool)
        -> Result<Vec<f32>, TruenoError> {
        let mut out = vec![0.0f32; heads * seq_len * head_dim];
        check(unsafe {
            sys::trn_attention_f32(q.as_ptr(), q.len(), k.as_ptr(), k.len(), v.as_ptr(), v.len(), out.as_mut_ptr(), heads, seq_len,
                                   head_dim, scale, causal as i32)
        })?;
        Ok(out)
    }
    pub fn matvec(a: &[f32], rows: usize, cols: usize, v: &[f32]) -> Result<Vec<f32>, TruenoError> {
        let mut y = vec![0.0f32; rows];
        check(unsafe { sys::trn_matvec_f32(a.as_ptr(), rows, cols, v.as_ptr(), v.len(), y.as_mut_ptr()) })?;
        Ok(y)
    }
    /// `Matrix::embedding_lookup` (src/matrix.rs:2008): rows of `table` selected by `indices`.
    pub fn embedding_lookup(table: &[f32], rows: usize, cols: usize, indices: &[usize]) -> Result<Vec<f32>, TruenoError> {
        let idx: Vec<u64> = indices.iter().map(|&i| i as u64).collect();
        let mut out = vec![0.0f32; indices.len() * cols];
        check(unsafe { sys::trn_embedding_lookup_f32(table.as_ptr(), rows, cols, idx.as_ptr(), idx.len(), out.as_mut_ptr()) })?;
        Ok(out)
    }
    pub fn transpose(a: &[f32], rows: usize, cols: usize) -> Result<Vec<f32>, TruenoError> {
        let mut out = vec![0.0f32; rows * cols];
        check(unsafe { sys::trn_transpose_f32(a.as_ptr(), rows, cols, out.as_mut_ptr()) })?;
        Ok(out)
    }
}
