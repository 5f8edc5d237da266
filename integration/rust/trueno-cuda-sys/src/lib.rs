//! Raw bindings to include/trueno_cuda.h — one declaration per exported symbol, nothing else.
//! Status codes: 0 OK, 1 SizeMismatch, 2 InvalidInput, 3 EmptyVector, 4 DivisionByZero,
//! 5 GpuError, 6 UnsupportedBackend (the variants of `TruenoError`, src/error.rs:8-41).
#![allow(non_camel_case_types)]
use core::ffi::{c_char, c_int, c_void};

#[repr(C)]
pub struct trn_buf {
    _private: [u8; 0],
}

/// (value, global index) of one slice of a sharded vector; index == u64::MAX means "no candidate"
#[repr(C)]
#[derive(Clone, Copy)]
pub struct trn_arg_pair {
    pub value: f32,
    pub reserved: u32,
    pub index: u64,
}

#[repr(C)]
pub struct trn_comm {
    _private: [u8; 0],
}

#[repr(C)]
pub struct trn_batch {
    _private: [u8; 0],
}

#[repr(C)]
pub struct trn_gemm_b {
    _private: [u8; 0],
}

pub const TRN_OK: c_int = 0;
pub const TRN_SIZE_MISMATCH: c_int = 1;
pub const TRN_INVALID_INPUT: c_int = 2;
pub const TRN_EMPTY_VECTOR: c_int = 3;
pub const TRN_DIVISION_BY_ZERO: c_int = 4;
pub const TRN_GPU_ERROR: c_int = 5;
pub const TRN_UNSUPPORTED_BACKEND: c_int = 6;

extern "C" {
    // context
    pub fn trn_cuda_init(device: c_int) -> c_int;
    pub fn trn_cuda_shutdown() -> c_int;
    pub fn trn_cuda_is_available() -> c_int;
    pub fn trn_device_count(count: *mut c_int) -> c_int;
    pub fn trn_device_info(name: *mut c_char, cap: usize, sm_count: *mut c_int, hbm_bytes: *mut u64) -> c_int;
    pub fn trn_last_error(buf: *mut c_char, cap: usize) -> usize;
    pub fn trn_last_mismatch(expected: *mut u64, actual: *mut u64);
    pub fn trn_synchronize(stream: *mut c_void) -> c_int;
    pub fn trn_launch_count() -> u64;
    // device buffers + pinned host memory
    pub fn trn_buf_alloc(len: usize, out: *mut *mut trn_buf) -> c_int;
    pub fn trn_buf_free(buf: *mut trn_buf) -> c_int;
    pub fn trn_buf_upload(buf: *mut trn_buf, host: *const f32, len: usize) -> c_int;
    pub fn trn_buf_download(buf: *const trn_buf, host: *mut f32, len: usize) -> c_int;
    pub fn trn_buf_len(buf: *const trn_buf) -> usize;
    pub fn trn_buf_ptr(buf: *const trn_buf) -> *mut f32;
    pub fn trn_host_alloc(len: usize, out: *mut *mut f32) -> c_int;
    pub fn trn_host_free(ptr: *mut f32) -> c_int;
    // host-slice operators (trait VectorBackend, src/backends/mod.rs:52-385; GpuBackend::matmul, gpu/mod.rs:434)
    pub fn trn_dot_f32(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32) -> c_int;
    pub fn trn_sum_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_max_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_min_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_argmax_f32(a: *const f32, n: usize, out: *mut u64) -> c_int;
    pub fn trn_argmin_f32(a: *const f32, n: usize, out: *mut u64) -> c_int;
    pub fn trn_norm_l2_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_add_f32(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32) -> c_int;
    pub fn trn_mul_f32(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32) -> c_int;
    pub fn trn_sigmoid_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_gelu_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_softmax_rows_f32(a: *const f32, out: *mut f32, rows: usize, cols: usize) -> c_int;
    pub fn trn_log_softmax_rows_f32(a: *const f32, out: *mut f32, rows: usize, cols: usize) -> c_int;
    pub fn trn_matmul_f32(a: *const f32, a_rows: usize, a_cols: usize, b: *const f32, b_rows: usize, b_cols: usize,
                          c: *mut f32) -> c_int;
    pub fn trn_batched_matmul_f32(a: *const f32, a_len: usize, b: *const f32, b_len: usize, c: *mut f32,
                                  batch: usize, m: usize, k: usize, n: usize) -> c_int;
    pub fn trn_batched_matmul_4d_f32(a: *const f32, a_len: usize, b: *const f32, b_len: usize, c: *mut f32,
                                     batch: usize, heads: usize, m: usize, k: usize, n: usize) -> c_int;
    pub fn trn_matvec_f32(a: *const f32, rows: usize, cols: usize, v: *const f32, v_len: usize, y: *mut f32) -> c_int;
    pub fn trn_transpose_f32(a: *const f32, rows: usize, cols: usize, out: *mut f32) -> c_int;
    // device-resident twins (stream-ordered; scalar outputs are device pointers)
    pub fn trn_dot_f32_dev(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_sum_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_max_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_min_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_argmax_f32_dev(a: *const f32, n: usize, out: *mut u64, out_value: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_argmin_f32_dev(a: *const f32, n: usize, out: *mut u64, out_value: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_argmax_slice_f32_dev(a: *const f32, n: usize, first_slice: c_int, out: *mut u64, out_value: *mut f32,
                                    stream: *mut c_void) -> c_int;
    pub fn trn_argmin_slice_f32_dev(a: *const f32, n: usize, first_slice: c_int, out: *mut u64, out_value: *mut f32,
                                    stream: *mut c_void) -> c_int;
    pub fn trn_argmax_slice_pair_f32_dev(a: *const f32, n: usize, slice_start: u64, out: *mut trn_arg_pair, stream: *mut c_void) -> c_int;
    pub fn trn_argmin_slice_pair_f32_dev(a: *const f32, n: usize, slice_start: u64, out: *mut trn_arg_pair, stream: *mut c_void) -> c_int;
    pub fn trn_arg_combine_f32_dev(pairs: *const trn_arg_pair, count: usize, is_max: c_int, out_idx: *mut u64, out_value: *mut f32,
                                   stream: *mut c_void) -> c_int;
    pub fn trn_norm_l2_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_sumsq_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_add_f32_dev(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_mul_f32_dev(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_sigmoid_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_gelu_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_softmax_rows_f32_dev(a: *const f32, out: *mut f32, rows: usize, cols: usize, stream: *mut c_void) -> c_int;
    pub fn trn_log_softmax_rows_f32_dev(a: *const f32, out: *mut f32, rows: usize, cols: usize, stream: *mut c_void) -> c_int;
    pub fn trn_matmul_f32_dev(a: *const f32, a_rows: usize, a_cols: usize, b: *const f32, b_rows: usize, b_cols: usize,
                              c: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_batched_matmul_f32_dev(a: *const f32, a_len: usize, b: *const f32, b_len: usize, c: *mut f32,
                                      batch: usize, m: usize, k: usize, n: usize, stream: *mut c_void) -> c_int;
    pub fn trn_batched_matmul_4d_f32_dev(a: *const f32, a_len: usize, b: *const f32, b_len: usize, c: *mut f32,
                                         batch: usize, heads: usize, m: usize, k: usize, n: usize,
                                         stream: *mut c_void) -> c_int;
    pub fn trn_matvec_f32_dev(a: *const f32, rows: usize, cols: usize, v: *const f32, v_len: usize, y: *mut f32,
                              stream: *mut c_void) -> c_int;
    pub fn trn_transpose_f32_dev(a: *const f32, rows: usize, cols: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    // remaining VectorBackend surface (src/backends/mod.rs:67-385)
    pub fn trn_abs_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_abs_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_relu_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_relu_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_exp_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_exp_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_swish_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_swish_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_tanh_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_tanh_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_sqrt_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_sqrt_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_recip_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_recip_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_ln_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_ln_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_log2_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_log2_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_log10_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_log10_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_sin_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_sin_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_cos_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_cos_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_tan_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_tan_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_floor_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_floor_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_ceil_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_ceil_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_round_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_round_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_sum_kahan_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_sum_kahan_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_norm_l1_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_norm_l1_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_norm_linf_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_norm_linf_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_sub_f32(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32) -> c_int;
    pub fn trn_sub_f32_dev(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_div_f32(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32) -> c_int;
    pub fn trn_div_f32_dev(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_scale_f32(a: *const f32, n: usize, scalar: f32, out: *mut f32) -> c_int;
    pub fn trn_scale_f32_dev(a: *const f32, n: usize, scalar: f32, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_clamp_f32(a: *const f32, n: usize, min_val: f32, max_val: f32, out: *mut f32) -> c_int;
    pub fn trn_clamp_f32_dev(a: *const f32, n: usize, min_val: f32, max_val: f32, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_lerp_f32(a: *const f32, na: usize, b: *const f32, nb: usize, t: f32, out: *mut f32) -> c_int;
    pub fn trn_lerp_f32_dev(a: *const f32, na: usize, b: *const f32, nb: usize, t: f32, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_fma_f32(a: *const f32, na: usize, b: *const f32, nb: usize, c: *const f32, nc: usize, out: *mut f32) -> c_int;
    pub fn trn_fma_f32_dev(a: *const f32, na: usize, b: *const f32, nb: usize, c: *const f32, nc: usize, out: *mut f32,
                           stream: *mut c_void) -> c_int;
    // the rest of Vector's element-wise / statistics API (include/trueno_cuda.h, same order)
    pub fn trn_neg_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_neg_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_signum_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_signum_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_trunc_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_trunc_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_fract_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_fract_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_sinh_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_sinh_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_cosh_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_cosh_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_asin_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_asin_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_acos_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_acos_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_atan_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_atan_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_asinh_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_asinh_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_acosh_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_acosh_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_atanh_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_atanh_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_hardswish_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_hardswish_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_mish_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_mish_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_selu_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_selu_f32_dev(a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_leaky_relu_f32(a: *const f32, n: usize, negative_slope: f32, out: *mut f32) -> c_int;
    pub fn trn_leaky_relu_f32_dev(a: *const f32, n: usize, negative_slope: f32, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_elu_f32(a: *const f32, n: usize, alpha: f32, out: *mut f32) -> c_int;
    pub fn trn_elu_f32_dev(a: *const f32, n: usize, alpha: f32, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_pow_f32(a: *const f32, n: usize, exponent: f32, out: *mut f32) -> c_int;
    pub fn trn_pow_f32_dev(a: *const f32, n: usize, exponent: f32, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_clip_f32(a: *const f32, n: usize, min_val: f32, max_val: f32, out: *mut f32) -> c_int;
    pub fn trn_clip_f32_dev(a: *const f32, n: usize, min_val: f32, max_val: f32, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_minimum_f32(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32) -> c_int;
    pub fn trn_minimum_f32_dev(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_maximum_f32(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32) -> c_int;
    pub fn trn_maximum_f32_dev(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_copysign_f32(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32) -> c_int;
    pub fn trn_copysign_f32_dev(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_affine_f32_dev(a: *const f32, n: usize, shift: f32, scale: f32, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_sum_of_squares_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_covariance_f32(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32) -> c_int;
    pub fn trn_correlation_f32(a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32) -> c_int;
    pub fn trn_zscore_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_minmax_normalize_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_layer_norm_simple_rows_f32(a: *const f32, eps: f32, out: *mut f32, rows: usize, cols: usize) -> c_int;
    pub fn trn_layer_norm_simple_rows_f32_dev(a: *const f32, eps: f32, out: *mut f32, rows: usize, cols: usize, stream: *mut c_void) -> c_int;
    pub fn trn_softmax_slice_stats_f32_dev(a: *const f32, n: usize, pair_out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_softmax_slice_apply_f32_dev(a: *const f32, n: usize, pairs: *const f32, npairs: usize, log_variant: c_int,
                                           out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_embedding_lookup_f32(table: *const f32, rows: usize, cols: usize, indices: *const u64, n_indices: usize,
                                    out: *mut f32) -> c_int;
    pub fn trn_embedding_lookup_f32_dev(table: *const f32, rows: usize, cols: usize, indices: *const u64, n_indices: usize,
                                        out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_mean_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_variance_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_stddev_f32(a: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn trn_vecmat_f32(v: *const f32, v_len: usize, a: *const f32, rows: usize, cols: usize, y: *mut f32) -> c_int;
    pub fn trn_vecmat_f32_dev(v: *const f32, v_len: usize, a: *const f32, rows: usize, cols: usize, y: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_layer_norm_rows_f32(a: *const f32, gamma: *const f32, gamma_len: usize, beta: *const f32, beta_len: usize,
                                   eps: f32, out: *mut f32, rows: usize, cols: usize) -> c_int;
    pub fn trn_layer_norm_rows_f32_dev(a: *const f32, gamma: *const f32, gamma_len: usize, beta: *const f32, beta_len: usize,
                                       eps: f32, out: *mut f32, rows: usize, cols: usize, stream: *mut c_void) -> c_int;
    pub fn trn_convolve2d_f32(input: *const f32, rows: usize, cols: usize, kernel: *const f32, k_rows: usize, k_cols: usize,
                              out: *mut f32) -> c_int;
    pub fn trn_convolve2d_f32_dev(input: *const f32, rows: usize, cols: usize, kernel: *const f32, k_rows: usize, k_cols: usize,
                                  out: *mut f32, stream: *mut c_void) -> c_int;
    // SymmetricEigen (src/eigen.rs:108): eigenvalues descending, eigenvectors as columns
    pub fn trn_symmetric_eigen_f32(a: *const f32, rows: usize, cols: usize, eigenvalues: *mut f32, eigenvectors: *mut f32) -> c_int;
    pub fn trn_symmetric_eigen_f32_dev(a: *const f32, rows: usize, cols: usize, eigenvalues: *mut f32, eigenvectors: *mut f32,
                                       stream: *mut c_void) -> c_int;
    // fused attention (trueno-gpu AttentionKernel): q, k, v, out are [heads][seq_len][head_dim]
    pub fn trn_attention_f32(q: *const f32, q_len: usize, k: *const f32, k_len: usize, v: *const f32, v_len: usize, out: *mut f32,
                             heads: usize, seq_len: usize, head_dim: usize, scale: f32, causal: c_int) -> c_int;
    pub fn trn_attention_f32_dev(q: *const f32, q_len: usize, k: *const f32, k_len: usize, v: *const f32, v_len: usize,
                                 out: *mut f32, heads: usize, seq_len: usize, head_dim: usize, scale: f32, causal: c_int,
                                 stream: *mut c_void) -> c_int;
    // fused slice reduction + exchange over NVLink peer memory
    pub fn trn_comm_local_handle(handle64: *mut c_void) -> c_int;
    pub fn trn_comm_create(rank: c_int, world: c_int, handles: *const c_void, out: *mut *mut trn_comm) -> c_int;
    pub fn trn_comm_destroy(comm: *mut trn_comm) -> c_int;
    pub fn trn_sum_allreduce_f32_dev(comm: *mut trn_comm, a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_dot_allreduce_f32_dev(comm: *mut trn_comm, a: *const f32, na: usize, b: *const f32, nb: usize, out: *mut f32,
                                     stream: *mut c_void) -> c_int;
    pub fn trn_norm_l2_allreduce_f32_dev(comm: *mut trn_comm, a: *const f32, n: usize, out: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_argmax_allgather_f32_dev(comm: *mut trn_comm, a: *const f32, n: usize, slice_start: u64, out_idx: *mut u64,
                                        out_value: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_argmin_allgather_f32_dev(comm: *mut trn_comm, a: *const f32, n: usize, slice_start: u64, out_idx: *mut u64,
                                        out_value: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_comm_status(comm: *mut trn_comm) -> c_int;
    // pre-split right-hand operand: many products against one B (row-block shards, SymmetricEigen::reconstruct, weights)
    pub fn trn_gemm_prepare_b_dev(b: *const f32, b_rows: usize, b_cols: usize, out: *mut *mut trn_gemm_b, stream: *mut c_void) -> c_int;
    pub fn trn_gemm_b_free(handle: *mut trn_gemm_b) -> c_int;
    pub fn trn_matmul_prepared_f32_dev(a: *const f32, a_rows: usize, a_cols: usize, b: *const trn_gemm_b, c: *mut f32,
                                       stream: *mut c_void) -> c_int;
    pub fn trn_matmul_prepared_f32(a: *const f32, a_rows: usize, a_cols: usize, b: *const trn_gemm_b, c: *mut f32) -> c_int;
    // row blocks of one sharded product: kernel chosen as for the whole total_rows-row product (bit-identical gather)
    pub fn trn_matmul_rowblock_f32_dev(a: *const f32, block_rows: usize, total_rows: usize, a_cols: usize, b: *const f32,
                                       b_rows: usize, b_cols: usize, c: *mut f32, stream: *mut c_void) -> c_int;
    pub fn trn_matmul_rowblock_prepared_f32_dev(a: *const f32, block_rows: usize, total_rows: usize, a_cols: usize,
                                                b: *const trn_gemm_b, c: *mut f32, stream: *mut c_void) -> c_int;
    // device-resident op chaining (GpuCommandBatch counterpart, src/backends/gpu/batch.rs)
    pub fn trn_batch_create(out: *mut *mut trn_batch) -> c_int;
    pub fn trn_batch_destroy(batch: *mut trn_batch) -> c_int;
    pub fn trn_batch_upload(batch: *mut trn_batch, data: *const f32, len: usize, id: *mut u32) -> c_int;
    pub fn trn_batch_update(batch: *mut trn_batch, id: u32, data: *const f32, len: usize) -> c_int;
    pub fn trn_batch_op(batch: *mut trn_batch, op: c_int, a_id: u32, b_id: u32, scalar: f32, out_id: *mut u32) -> c_int;
    pub fn trn_batch_execute(batch: *mut trn_batch) -> c_int;
    pub fn trn_batch_read(batch: *mut trn_batch, id: u32, out: *mut f32, len: usize) -> c_int;
    pub fn trn_batch_num_operations(batch: *const trn_batch) -> usize;
    pub fn trn_batch_num_buffers(batch: *const trn_batch) -> usize;
    // engine selection / live timing (benches only)
    pub fn trn_set_gemm_engine(engine: c_int) -> c_int;
    pub fn trn_get_gemm_engine() -> c_int;
    pub fn trn_profile_enable(on: c_int) -> c_int;
    pub fn trn_profile_last_gemm(prepass_ms: *mut f32, kernel_ms: *mut f32) -> c_int;
}
