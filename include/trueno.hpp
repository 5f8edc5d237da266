// trueno.hpp — C++17 host mirror of trueno's public API for the hot path, over the C ABI (trueno_cuda.h).
//
// The reference is a Rust crate; where no Rust toolchain exists this header is the compiled-language host
// side: the same names, argument meaning and error behaviour as `Vector<f32>` (src/vector.rs), `Matrix<f32>`
// (src/matrix.rs), `TruenoError` / `Result` (src/error.rs:8-41) and `GpuCommandBatch`
// (src/backends/gpu/batch.rs), so tests written against it read like the reference's own tests
// (tests/cpp/test_trueno_hpp.cpp).  Header-only; link with -ltrueno_cuda.  Every operation runs on the B200
// through the C ABI — there is no CPU fallback here either: a missing device surfaces as TruenoError::GpuError.
#pragma once

#include <cmath>
#include <optional>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <utility>
#include <variant>
#include <vector>

#include "trueno_cuda.h"

namespace trueno {

// ---- TruenoError (src/error.rs:8-41) -------------------------------------------------------------------
struct TruenoError {
    enum Kind { UnsupportedBackend, SizeMismatch, GpuError, InvalidInput, DivisionByZero, EmptyVector };
    Kind kind = GpuError;
    std::string message;            // payload of InvalidInput / GpuError / UnsupportedBackend
    size_t expected = 0, actual = 0;  // payload of SizeMismatch

    static TruenoError size_mismatch(size_t e, size_t a) { TruenoError x; x.kind = SizeMismatch; x.expected = e; x.actual = a; return x; }
    static TruenoError invalid_input(std::string m) { TruenoError x; x.kind = InvalidInput; x.message = std::move(m); return x; }
    static TruenoError empty_vector() { TruenoError x; x.kind = EmptyVector; return x; }
    static TruenoError division_by_zero() { TruenoError x; x.kind = DivisionByZero; return x; }

    // Display text, byte-identical to the #[error(...)] strings of src/error.rs
    std::string to_string() const {
        switch (kind) {
            case UnsupportedBackend: return "Backend not supported on this platform: " + message;
            case SizeMismatch: return "Size mismatch: expected " + std::to_string(expected) + ", got " + std::to_string(actual);
            case GpuError: return "GPU error: " + message;
            case InvalidInput: return "Invalid input: " + message;
            case DivisionByZero: return "Division by zero";
            case EmptyVector: return "Empty vector";
        }
        return {};
    }
    // derive(PartialEq): variant and payload
    bool operator==(const TruenoError& o) const {
        if (kind != o.kind) return false;
        if (kind == SizeMismatch) return expected == o.expected && actual == o.actual;
        if (kind == EmptyVector || kind == DivisionByZero) return true;
        return message == o.message;
    }
    bool operator!=(const TruenoError& o) const { return !(*this == o); }
};

// ---- Result<T> (src/error.rs:8) ------------------------------------------------------------------------
template <class T>
class Result {
    std::variant<T, TruenoError> v_;

public:
    Result(T value) : v_(std::move(value)) {}
    Result(TruenoError e) : v_(std::move(e)) {}
    bool is_ok() const { return v_.index() == 0; }
    bool is_err() const { return v_.index() == 1; }
    // like Rust's unwrap(): a programming error aborts with the error text
    T& unwrap() {
        if (is_err()) { fprintf(stderr, "called `Result::unwrap()` on an `Err` value: %s\n", std::get<1>(v_).to_string().c_str()); std::abort(); }
        return std::get<0>(v_);
    }
    const T& unwrap() const { return const_cast<Result*>(this)->unwrap(); }
    const TruenoError& unwrap_err() const {
        if (is_ok()) { fprintf(stderr, "called `Result::unwrap_err()` on an `Ok` value\n"); std::abort(); }
        return std::get<1>(v_);
    }
};

namespace detail {
inline TruenoError from_status(int status) {
    char buf[1024];
    trn_last_error(buf, sizeof buf);
    TruenoError e;
    e.message = buf;
    switch (status) {
        case TRN_SIZE_MISMATCH: {
            uint64_t ex = 0, ac = 0;
            trn_last_mismatch(&ex, &ac);
            return TruenoError::size_mismatch((size_t)ex, (size_t)ac);
        }
        case TRN_INVALID_INPUT: e.kind = TruenoError::InvalidInput; break;
        case TRN_EMPTY_VECTOR: return TruenoError::empty_vector();
        case TRN_DIVISION_BY_ZERO: return TruenoError::division_by_zero();
        case TRN_UNSUPPORTED_BACKEND: e.kind = TruenoError::UnsupportedBackend; break;
        default: e.kind = TruenoError::GpuError; break;
    }
    return e;
}
}  // namespace detail

class Matrix;

// ---- Vector<f32> (src/vector.rs:125-128) ----------------------------------------------------------------
class Vector {
    std::vector<float> data_;

    template <class F>
    Result<float> reduce(F fn) const {
        float out = 0.f;
        const int st = fn(data_.data(), data_.size(), &out);
        if (st != TRN_OK) return detail::from_status(st);
        return out;
    }
    template <class F>
    Result<Vector> map(F fn) const {
        Vector out{std::vector<float>(data_.size())};
        const int st = fn(data_.data(), data_.size(), out.data_.data());
        if (st != TRN_OK) return detail::from_status(st);
        return out;
    }
    template <class F>
    Result<Vector> zip(const Vector& o, F fn) const {
        Vector out{std::vector<float>(data_.size() == o.data_.size() ? data_.size() : 0)};
        const int st = fn(data_.data(), data_.size(), o.data_.data(), o.data_.size(), out.data_.data());
        if (st != TRN_OK) return detail::from_status(st);
        return out;
    }

public:
    Vector() = default;
    explicit Vector(std::vector<float> d) : data_(std::move(d)) {}
    static Vector from_slice(const std::vector<float>& d) { return Vector(d); }            // src/vector.rs:151
    static Vector from_slice(const float* p, size_t n) { return Vector(std::vector<float>(p, p + n)); }
    static Vector from_vec(std::vector<float> d) { return Vector(std::move(d)); }          // src/vector.rs:199
    size_t len() const { return data_.size(); }
    bool is_empty() const { return data_.empty(); }
    const std::vector<float>& as_slice() const { return data_; }                             // src/vector.rs:293
    bool operator==(const Vector& o) const { return data_ == o.data_; }

    // reductions (src/vector.rs:588-827, 848, 2601-2770)
    Result<float> dot(const Vector& o) const {
        float out = 0.f;
        const int st = trn_dot_f32(data_.data(), data_.size(), o.data_.data(), o.data_.size(), &out);
        if (st != TRN_OK) return detail::from_status(st);
        return out;
    }
    Result<float> sum() const { return reduce(trn_sum_f32); }
    Result<float> sum_kahan() const { return reduce(trn_sum_kahan_f32); }
    Result<float> max() const { return reduce(trn_max_f32); }
    Result<float> min() const { return reduce(trn_min_f32); }
    Result<float> norm_l2() const { return reduce(trn_norm_l2_f32); }
    Result<float> norm_l1() const { return reduce(trn_norm_l1_f32); }
    Result<float> norm_linf() const { return reduce(trn_norm_linf_f32); }
    Result<float> mean() const { return reduce(trn_mean_f32); }
    Result<float> variance() const { return reduce(trn_variance_f32); }
    Result<float> stddev() const { return reduce(trn_stddev_f32); }
    Result<size_t> argmax() const {
        uint64_t out = 0;
        const int st = trn_argmax_f32(data_.data(), data_.size(), &out);
        if (st != TRN_OK) return detail::from_status(st);
        return (size_t)out;
    }
    Result<size_t> argmin() const {
        uint64_t out = 0;
        const int st = trn_argmin_f32(data_.data(), data_.size(), &out);
        if (st != TRN_OK) return detail::from_status(st);
        return (size_t)out;
    }

    // elementwise (src/vector.rs:358-533, 1670-4182)
    Result<Vector> add(const Vector& o) const { return zip(o, trn_add_f32); }
    Result<Vector> sub(const Vector& o) const { return zip(o, trn_sub_f32); }
    Result<Vector> mul(const Vector& o) const { return zip(o, trn_mul_f32); }
    Result<Vector> div(const Vector& o) const { return zip(o, trn_div_f32); }
    Result<Vector> abs() const { return map(trn_abs_f32); }
    Result<Vector> relu() const { return map(trn_relu_f32); }
    Result<Vector> exp() const { return map(trn_exp_f32); }
    Result<Vector> sigmoid() const { return map(trn_sigmoid_f32); }
    Result<Vector> gelu() const { return map(trn_gelu_f32); }
    Result<Vector> swish() const { return map(trn_swish_f32); }
    Result<Vector> tanh() const { return map(trn_tanh_f32); }
    Result<Vector> sqrt() const { return map(trn_sqrt_f32); }
    Result<Vector> recip() const { return map(trn_recip_f32); }
    Result<Vector> ln() const { return map(trn_ln_f32); }
    Result<Vector> log2() const { return map(trn_log2_f32); }
    Result<Vector> log10() const { return map(trn_log10_f32); }
    Result<Vector> sin() const { return map(trn_sin_f32); }
    Result<Vector> cos() const { return map(trn_cos_f32); }
    Result<Vector> tan() const { return map(trn_tan_f32); }
    Result<Vector> floor() const { return map(trn_floor_f32); }
    Result<Vector> ceil() const { return map(trn_ceil_f32); }
    Result<Vector> round() const { return map(trn_round_f32); }
    Result<Vector> scale(float s) const {
        return map([s](const float* a, size_t n, float* o) { return trn_scale_f32(a, n, s, o); });
    }
    Result<Vector> clamp(float lo, float hi) const {
        return map([lo, hi](const float* a, size_t n, float* o) { return trn_clamp_f32(a, n, lo, hi, o); });
    }
    Result<Vector> lerp(const Vector& o, float t) const {
        return zip(o, [t](const float* a, size_t na, const float* b, size_t nb, float* out) { return trn_lerp_f32(a, na, b, nb, t, out); });
    }
    Result<Vector> fma(const Vector& b, const Vector& c) const {
        Vector out{std::vector<float>(data_.size())};
        const int st = trn_fma_f32(data_.data(), data_.size(), b.data_.data(), b.data_.size(), c.data_.data(), c.data_.size(), out.data_.data());
        if (st != TRN_OK) return detail::from_status(st);
        return out;
    }
    // the rest of Vector's element-wise / statistics API (src/vector.rs:898-1480, 1980-2570, 3342-4410)
    Result<Vector> neg() const { return map(trn_neg_f32); }
    Result<Vector> signum() const { return map(trn_signum_f32); }
    Result<Vector> trunc() const { return map(trn_trunc_f32); }
    Result<Vector> fract() const { return map(trn_fract_f32); }
    Result<Vector> sinh() const { return map(trn_sinh_f32); }
    Result<Vector> cosh() const { return map(trn_cosh_f32); }
    Result<Vector> asin() const { return map(trn_asin_f32); }
    Result<Vector> acos() const { return map(trn_acos_f32); }
    Result<Vector> atan() const { return map(trn_atan_f32); }
    Result<Vector> asinh() const { return map(trn_asinh_f32); }
    Result<Vector> acosh() const { return map(trn_acosh_f32); }
    Result<Vector> atanh() const { return map(trn_atanh_f32); }
    Result<Vector> hardswish() const { return map(trn_hardswish_f32); }
    Result<Vector> mish() const { return map(trn_mish_f32); }
    Result<Vector> selu() const { return map(trn_selu_f32); }
    Result<Vector> leaky_relu(float negative_slope) const {
        return map([negative_slope](const float* a, size_t n, float* o) { return trn_leaky_relu_f32(a, n, negative_slope, o); });
    }
    Result<Vector> elu(float alpha) const {
        return map([alpha](const float* a, size_t n, float* o) { return trn_elu_f32(a, n, alpha, o); });
    }
    Result<Vector> pow(float e) const {
        return map([e](const float* a, size_t n, float* o) { return trn_pow_f32(a, n, e, o); });
    }
    Result<Vector> clip(float lo, float hi) const {
        return map([lo, hi](const float* a, size_t n, float* o) { return trn_clip_f32(a, n, lo, hi, o); });
    }
    Result<Vector> minimum(const Vector& o) const { return zip(o, trn_minimum_f32); }
    Result<Vector> maximum(const Vector& o) const { return zip(o, trn_maximum_f32); }
    Result<Vector> copysign(const Vector& sign) const { return zip(sign, trn_copysign_f32); }
    Result<Vector> zscore() const { return map(trn_zscore_f32); }
    Result<Vector> minmax_normalize() const { return map(trn_minmax_normalize_f32); }
    Result<Vector> layer_norm_simple(float eps) const {
        return map([eps](const float* a, size_t n, float* o) { return trn_layer_norm_simple_rows_f32(a, eps, o, 1, n); });
    }
    Result<float> sum_of_squares() const { return reduce(trn_sum_of_squares_f32); }
    Result<float> covariance(const Vector& o) const {
        float out = 0.f;
        const int st = trn_covariance_f32(data_.data(), data_.size(), o.data_.data(), o.data_.size(), &out);
        if (st != TRN_OK) return detail::from_status(st);
        return out;
    }
    Result<float> correlation(const Vector& o) const {
        float out = 0.f;
        const int st = trn_correlation_f32(data_.data(), data_.size(), o.data_.data(), o.data_.size(), &out);
        if (st != TRN_OK) return detail::from_status(st);
        return out;
    }
    Result<Vector> normalize() const {                                                      // src/vector.rs:2665-2678
        auto n = norm_l2();
        if (n.is_err()) return n.unwrap_err();
        if (std::abs(n.unwrap()) < 1e-10f) return TruenoError::division_by_zero();
        return scale(1.0f / n.unwrap());
    }
    Result<Vector> layer_norm(const Vector& gamma, const Vector& beta, float eps) const {   // src/vector.rs:1316
        Vector out{std::vector<float>(data_.size())};
        const int st = trn_layer_norm_rows_f32(data_.data(), gamma.data_.data(), gamma.data_.size(), beta.data_.data(),
                                               beta.data_.size(), eps, out.data_.data(), data_.empty() ? 0 : 1, data_.size());
        if (st != TRN_OK) return detail::from_status(st);
        return out;
    }
    // softmax family (src/vector.rs:1516, 1581): a Vector is one row
    Result<Vector> softmax() const {
        return map([](const float* a, size_t n, float* o) { return trn_softmax_rows_f32(a, o, n ? 1 : 0, n); });
    }
    Result<Vector> log_softmax() const {
        return map([](const float* a, size_t n, float* o) { return trn_log_softmax_rows_f32(a, o, n ? 1 : 0, n); });
    }
    friend class Matrix;
};

// ---- Matrix<f32> (src/matrix.rs:49-54): dense row-major ---------------------------------------------------
class Matrix {
    size_t rows_ = 0, cols_ = 0;
    std::vector<float> data_;
    Matrix(size_t r, size_t c, std::vector<float> d) : rows_(r), cols_(c), data_(std::move(d)) {}

public:
    static Result<Matrix> from_vec(size_t rows, size_t cols, std::vector<float> data) {   // src/matrix.rs:108-117
        if (data.size() != rows * cols)
            return TruenoError::invalid_input("Data length " + std::to_string(data.size()) + " does not match matrix dimensions " +
                                              std::to_string(rows) + "x" + std::to_string(cols) + " (expected " +
                                              std::to_string(rows * cols) + ")");
        return Matrix(rows, cols, std::move(data));
    }
    static Matrix zeros(size_t rows, size_t cols) { return Matrix(rows, cols, std::vector<float>(rows * cols, 0.f)); }
    static Matrix identity(size_t n) {
        Matrix m = zeros(n, n);
        for (size_t i = 0; i < n; ++i) m.data_[i * n + i] = 1.f;
        return m;
    }
    size_t rows() const { return rows_; }
    size_t cols() const { return cols_; }
    const std::vector<float>& as_slice() const { return data_; }
    const float* get(size_t i, size_t j) const { return i < rows_ && j < cols_ ? &data_[i * cols_ + j] : nullptr; }

    Result<Matrix> matmul(const Matrix& o) const {                                         // src/matrix.rs:285
        Matrix out(rows_, o.cols_, std::vector<float>(cols_ == o.rows_ ? rows_ * o.cols_ : 0));
        const int st = trn_matmul_f32(data_.data(), rows_, cols_, o.data_.data(), o.rows_, o.cols_, out.data_.data());
        if (st != TRN_OK) return detail::from_status(st);
        return out;
    }
    Matrix transpose() const {                                                              // src/matrix.rs:1590
        Matrix out(cols_, rows_, std::vector<float>(data_.size()));
        trn_transpose_f32(data_.data(), rows_, cols_, out.data_.data());
        return out;
    }
    Result<Vector> matvec(const Vector& v) const {                                          // src/matrix.rs:1657
        Vector out{std::vector<float>(rows_)};
        const int st = trn_matvec_f32(data_.data(), rows_, cols_, v.data_.data(), v.data_.size(), out.data_.data());
        if (st != TRN_OK) return detail::from_status(st);
        return out;
    }
    Result<Matrix> convolve2d(const Matrix& k) const {                                       // src/matrix.rs:1868
        const bool ok = k.rows_ <= rows_ && k.cols_ <= cols_;
        Matrix out(ok ? rows_ - k.rows_ + 1 : 0, ok ? cols_ - k.cols_ + 1 : 0, {});
        out.data_.resize(out.rows_ * out.cols_);
        const int st = trn_convolve2d_f32(data_.data(), rows_, cols_, k.data_.data(), k.rows_, k.cols_, out.data_.data());
        if (st != TRN_OK) return detail::from_status(st);
        return out;
    }
    Result<Matrix> embedding_lookup(const std::vector<size_t>& indices) const {             // src/matrix.rs:2008
        std::vector<uint64_t> idx(indices.begin(), indices.end());
        Matrix out(indices.size(), cols_, {});
        out.data_.resize(out.rows_ * out.cols_);
        const int st = trn_embedding_lookup_f32(data_.data(), rows_, cols_, idx.data(), idx.size(), out.data_.data());
        if (st != TRN_OK) return detail::from_status(st);
        return out;
    }
    static Result<Vector> vecmat(const Vector& v, const Matrix& m) {                        // src/matrix.rs:1782
        Vector out{std::vector<float>(m.cols_)};
        const int st = trn_vecmat_f32(v.data_.data(), v.data_.size(), m.data_.data(), m.rows_, m.cols_, out.data_.data());
        if (st != TRN_OK) return detail::from_status(st);
        return out;
    }
    static Result<std::vector<float>> batched_matmul(const std::vector<float>& a, const std::vector<float>& b, size_t batch,
                                                     size_t m, size_t k, size_t n) {       // src/matrix.rs:383
        std::vector<float> c(batch * m * n);
        const int st = trn_batched_matmul_f32(a.data(), a.size(), b.data(), b.size(), c.data(), batch, m, k, n);
        if (st != TRN_OK) return detail::from_status(st);
        return c;
    }
    static Result<std::vector<float>> batched_matmul_4d(const std::vector<float>& a, const std::vector<float>& b, size_t batch,
                                                        size_t heads, size_t m, size_t k, size_t n) {   // src/matrix.rs:464
        std::vector<float> c(batch * heads * m * n);
        const int st = trn_batched_matmul_4d_f32(a.data(), a.size(), b.data(), b.size(), c.data(), batch, heads, m, k, n);
        if (st != TRN_OK) return detail::from_status(st);
        return c;
    }
};

// ---- SymmetricEigen (src/eigen.rs:56-516): eigenvalues descending, eigenvectors as the columns of a Matrix ----
class SymmetricEigen {
    std::vector<float> values_;
    Matrix vectors_;
    SymmetricEigen(std::vector<float> v, Matrix m) : values_(std::move(v)), vectors_(std::move(m)) {}

public:
    static Result<SymmetricEigen> create(const Matrix& m) {                                 // SymmetricEigen::new, src/eigen.rs:108
        const size_t n = m.rows() == m.cols() ? m.rows() : 0;
        std::vector<float> vals(n), vecs(n * n);
        const int st = trn_symmetric_eigen_f32(m.as_slice().data(), m.rows(), m.cols(), vals.data(), vecs.data());
        if (st != TRN_OK) return detail::from_status(st);
        auto mat = Matrix::from_vec(n, n, std::move(vecs));
        if (mat.is_err()) return mat.unwrap_err();
        return SymmetricEigen(std::move(vals), mat.unwrap());
    }
    const std::vector<float>& eigenvalues() const { return values_; }                       // src/eigen.rs:363
    const Matrix& eigenvectors() const { return vectors_; }                                 // src/eigen.rs:385
    size_t len() const { return values_.size(); }
    bool is_empty() const { return values_.empty(); }
    std::optional<Vector> eigenvector(size_t i) const {                                     // src/eigen.rs:434
        if (i >= len()) return std::nullopt;
        std::vector<float> col(len());
        for (size_t r = 0; r < len(); ++r) col[r] = vectors_.as_slice()[r * len() + i];
        return Vector{std::move(col)};
    }
};

// ---- fused attention: the counterpart of trueno-gpu's AttentionKernel (trueno-gpu/src/kernels/attention.rs:27-125) ----
// AttentionKernel::new(seq_len, head_dim) [.with_causal()] [.with_scale(s)]; run(q, k, v, heads) with q/k/v laid out
// [heads][seq_len][head_dim] returns softmax(scale * Q K^T [causal]) V per head.
class AttentionKernel {
    size_t seq_len_, head_dim_;
    float scale_;
    bool causal_ = false;

public:
    AttentionKernel(size_t seq_len, size_t head_dim)
        : seq_len_(seq_len), head_dim_(head_dim), scale_(1.0f / std::sqrt(static_cast<float>(head_dim))) {}
    AttentionKernel with_causal() const { AttentionKernel k = *this; k.causal_ = true; return k; }
    AttentionKernel with_scale(float s) const { AttentionKernel k = *this; k.scale_ = s; return k; }
    size_t seq_len() const { return seq_len_; }
    size_t head_dim() const { return head_dim_; }
    float scale() const { return scale_; }
    bool causal() const { return causal_; }
    Result<std::vector<float>> run(const std::vector<float>& q, const std::vector<float>& k, const std::vector<float>& v,
                                   size_t heads) const {
        std::vector<float> out(heads * seq_len_ * head_dim_);
        const int st = trn_attention_f32(q.data(), q.size(), k.data(), k.size(), v.data(), v.size(), out.data(), heads, seq_len_,
                                         head_dim_, scale_, causal_ ? 1 : 0);
        if (st != TRN_OK) return detail::from_status(st);
        return out;
    }
};

// ---- device-buffer type: f32 storage resident in HBM (trn_buf_*) --------------------------------------------
class DeviceBuffer {
    trn_buf* raw_ = nullptr;

public:
    DeviceBuffer() = default;
    DeviceBuffer(const DeviceBuffer&) = delete;
    DeviceBuffer& operator=(const DeviceBuffer&) = delete;
    DeviceBuffer(DeviceBuffer&& o) noexcept : raw_(o.raw_) { o.raw_ = nullptr; }
    ~DeviceBuffer() { if (raw_) trn_buf_free(raw_); }
    static Result<DeviceBuffer> with_len(size_t n) {
        DeviceBuffer b;
        const int st = trn_buf_alloc(n, &b.raw_);
        if (st != TRN_OK) return detail::from_status(st);
        return Result<DeviceBuffer>(std::move(b));
    }
    static Result<DeviceBuffer> from_slice(const std::vector<float>& d) {
        auto r = with_len(d.size());
        if (r.is_err()) return r.unwrap_err();
        const int st = trn_buf_upload(r.unwrap().raw_, d.data(), d.size());
        if (st != TRN_OK) return detail::from_status(st);
        return r;
    }
    size_t len() const { return trn_buf_len(raw_); }
    float* ptr() const { return trn_buf_ptr(raw_); }
    Result<std::vector<float>> to_vec() const {
        std::vector<float> out(len());
        const int st = trn_buf_download(raw_, out.data(), out.size());
        if (st != TRN_OK) return detail::from_status(st);
        return out;
    }
};

// ---- GpuCommandBatch (src/backends/gpu/batch.rs:118-1019) -------------------------------------------------
using BufferId = uint32_t;
class CommandBatch {
    trn_batch* raw_ = nullptr;
    std::vector<size_t> sizes_;
    BufferId op(int code, BufferId a, BufferId b = 0, float scalar = 0.f) {
        BufferId out = 0;
        if (trn_batch_op(raw_, code, a, b, scalar, &out) != TRN_OK) {   // the reference panics here (batch.rs:215-232)
            fprintf(stderr, "%s\n", detail::from_status(TRN_INVALID_INPUT).message.c_str());
            std::abort();
        }
        sizes_.push_back(code == 4 ? 1 : sizes_[a]);
        return out;
    }

public:
    CommandBatch() { trn_batch_create(&raw_); }
    CommandBatch(const CommandBatch&) = delete;
    CommandBatch& operator=(const CommandBatch&) = delete;
    ~CommandBatch() { trn_batch_destroy(raw_); }
    BufferId upload(const std::vector<float>& d) {
        BufferId id = 0;
        trn_batch_upload(raw_, d.data(), d.size(), &id);
        sizes_.push_back(d.size());
        return id;
    }
    BufferId relu(BufferId x) { return op(0, x); }
    BufferId scale(BufferId x, float s) { return op(1, x, 0, s); }
    BufferId add(BufferId a, BufferId b) { return op(2, a, b); }
    BufferId mul(BufferId a, BufferId b) { return op(3, a, b); }
    BufferId dot(BufferId a, BufferId b) { return op(4, a, b); }
    BufferId sigmoid(BufferId x) { return op(5, x); }
    BufferId tanh(BufferId x) { return op(6, x); }
    BufferId swish(BufferId x) { return op(7, x); }
    BufferId gelu(BufferId x) { return op(8, x); }
    BufferId sub(BufferId a, BufferId b) { return op(9, a, b); }
    Result<bool> execute() {
        const int st = trn_batch_execute(raw_);
        if (st != TRN_OK) return detail::from_status(st);
        return true;
    }
    Result<std::vector<float>> read(BufferId id) const {
        std::vector<float> out(id < sizes_.size() ? sizes_[id] : 0);
        const int st = trn_batch_read(raw_, id, out.data(), out.size());
        if (st != TRN_OK) return detail::from_status(st);
        return out;
    }
    size_t num_operations() const { return trn_batch_num_operations(raw_); }
    size_t num_buffers() const { return trn_batch_num_buffers(raw_); }
};

}  // namespace trueno
