// api.cu — the extern "C" operator surface (include/trueno_cuda.h): validation with the
// reference's exact error values/messages, engine dispatch, and the host-slice wrappers that
// stage through HBM.  No entry point ever computes on the CPU.
#include <atomic>
#include <cstdio>

#include "common.cuh"

using namespace trn;

namespace {

std::atomic<int> g_engine{0};

// U+00D7 MULTIPLICATION SIGN, as the reference's format strings use (src/matrix.rs:288, :398, :484)
#define X "\xC3\x97"

int need_ctx() { return ctx() ? TRN_OK : TRN_GPU_ERROR; }

// ---- validation shared by the host and device entry points ------------------------------------
int check_nonempty_invalid(size_t n) {  // src/vector.rs:654-656, 702-704, 750-752, 798-800
    return n == 0 ? fail(TRN_INVALID_INPUT, "Empty vector") : TRN_OK;
}
int check_nonempty_emptyvec(size_t n) {  // src/vector.rs:1517-1519, 1582-1584, 1855-1857, 2180-2182
    return n == 0 ? fail(TRN_EMPTY_VECTOR, "Empty vector") : TRN_OK;
}
int check_same_len(size_t na, size_t nb) {  // src/vector.rs:589-594, 359-364, 479-484
    return na != nb ? fail_mismatch(na, nb) : TRN_OK;
}
int check_matmul(size_t ar, size_t ac, size_t br, size_t bc) {  // src/matrix.rs:286-291
    if (ac != br)
        return fail(TRN_INVALID_INPUT,
                    "Matrix dimension mismatch for multiplication: %zu" X "%zu " X " %zu" X "%zu "
                    "(inner dimensions %zu and %zu must match)",
                    ar, ac, br, bc, ac, br);
    return TRN_OK;
}
int check_batched(size_t a_len, size_t b_len, size_t batch, size_t m, size_t k, size_t n) {  // src/matrix.rs:396-415
    if (a_len != batch * m * k)
        return fail(TRN_INVALID_INPUT, "A data size mismatch: expected %zu (%zu" X "%zu" X "%zu), got %zu",
                    batch * m * k, batch, m, k, a_len);
    if (b_len != batch * k * n)
        return fail(TRN_INVALID_INPUT, "B data size mismatch: expected %zu (%zu" X "%zu" X "%zu), got %zu",
                    batch * k * n, batch, k, n, b_len);
    return TRN_OK;
}
int check_batched_4d(size_t a_len, size_t b_len, size_t batch, size_t heads, size_t m, size_t k, size_t n) {
    const size_t total = batch * heads;  // src/matrix.rs:481-502
    if (a_len != total * m * k)
        return fail(TRN_INVALID_INPUT,
                    "A data size mismatch: expected %zu (%zu" X "%zu" X "%zu" X "%zu), got %zu", total * m * k,
                    batch, heads, m, k, a_len);
    if (b_len != total * k * n)
        return fail(TRN_INVALID_INPUT,
                    "B data size mismatch: expected %zu (%zu" X "%zu" X "%zu" X "%zu), got %zu", total * k * n,
                    batch, heads, k, n, b_len);
    return TRN_OK;
}
int check_matvec(size_t cols, size_t v_len) {  // src/matrix.rs:1658-1664
    if (v_len != cols)
        return fail(TRN_INVALID_INPUT,
                    "Vector length %zu does not match matrix columns %zu for matrix-vector multiplication", v_len,
                    cols);
    return TRN_OK;
}
#undef X

// ---- GEMM engine dispatch -----------------------------------------------------------------------
// Matrix::matmul routes by shape (src/matrix.rs:293-356); the CUDA backend routes every shape to
// the device (north star: unconditional, no thresholds that fall back to the CPU) and only picks
// WHICH device kernel: rows == 1 -> vecmat (reference quirk preserved); tiles that fill a tcgen05
// tile -> 3xTF32; the rest -> SIMT FFMA.
int gemm_dispatch(const float* a, const float* b, float* c, size_t batch, size_t m, size_t k, size_t n,
                  cudaStream_t s) {
    if (batch == 0 || m == 0 || n == 0) return TRN_OK;
    if (m == 1 && k > 0) {
        for (size_t i = 0; i < batch; ++i) TRN_TRY(launch_vecmat(a + i * k, b + i * k * n, k, n, c + i * n, s));
        return TRN_OK;
    }
    const int engine = g_engine.load();
    if (engine == 1) return launch_gemm_simt(a, b, c, batch, m, k, n, s);
    if (engine == 2 || engine == 3) {
        if (!gemm_tc_supported(m, k, n))
            return fail(TRN_INVALID_INPUT, "tcgen05 GEMM engine forced for an unsupported shape %zux%zux%zu", m, k, n);
        return launch_gemm_tc(a, b, c, batch, m, k, n, engine == 2 ? 3 : 1, s);
    }
    // auto: the tensor-core tile is 128x256; below that the SIMT kernel wins.  The choice depends on
    // (m, k, n) only — never on batch — so a batched product is bit-identical to the loop of single
    // products the reference runs (src/matrix.rs:507-524).
    if (gemm_tc_supported(m, k, n) && m >= 128 && n >= 128 && k >= 32 && m * n * k >= (size_t)1 << 24)
        return launch_gemm_tc(a, b, c, batch, m, k, n, 3, s);
    return launch_gemm_simt(a, b, c, batch, m, k, n, s);
}

// ---- host-slice plumbing ------------------------------------------------------------------------
struct DevTemp {
    float* p = nullptr;
    cudaStream_t s;
    explicit DevTemp(cudaStream_t st) : s(st) {}
    int alloc(size_t n) { return scratch_alloc((void**)&p, n * sizeof(float), s); }
    ~DevTemp() { if (p) cudaFreeAsync(p, s); }
};

template <class F>
int host_reduce_f32(const float* a, size_t na, const float* b, size_t nb, float* out, F&& launch) {
    Context* c = ctx();
    Workspace* w = workspace(c->stream);
    if (!w) return fail(TRN_GPU_ERROR, "failed to allocate the reduction workspace");
    DevTemp da(c->stream), db(c->stream);
    TRN_TRY(da.alloc(na));
    TRN_TRY(upload(da.p, a, na, c->stream));
    if (b) {
        TRN_TRY(db.alloc(nb));
        TRN_TRY(upload(db.p, b, nb, c->stream));
    }
    TRN_TRY(launch(da.p, db.p, w->scalar_f32, c->stream));
    TRN_CUDA(cudaMemcpyAsync(w->host_f32, w->scalar_f32, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    TRN_CUDA(cudaStreamSynchronize(c->stream));
    *out = *w->host_f32;
    return TRN_OK;
}

int host_arg(int is_max, const float* a, size_t n, uint64_t* out_idx, float* out_val) {
    Context* c = ctx();
    Workspace* w = workspace(c->stream);
    if (!w) return fail(TRN_GPU_ERROR, "failed to allocate the reduction workspace");
    DevTemp da(c->stream);
    TRN_TRY(da.alloc(n));
    TRN_TRY(upload(da.p, a, n, c->stream));
    TRN_TRY(launch_argreduce(is_max, da.p, n, w->scalar_u64, w->scalar_f32, c->stream));
    TRN_CUDA(cudaMemcpyAsync(w->host_u64, w->scalar_u64, sizeof(uint64_t), cudaMemcpyDeviceToHost, c->stream));
    TRN_CUDA(cudaMemcpyAsync(w->host_f32, w->scalar_f32, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    TRN_CUDA(cudaStreamSynchronize(c->stream));
    if (out_idx) *out_idx = *w->host_u64;
    if (out_val) *out_val = *w->host_f32;
    return TRN_OK;
}

int host_map(Map op, const float* a, const float* b, float* out, size_t n) {
    Context* c = ctx();
    DevTemp da(c->stream), db(c->stream), dout(c->stream);
    TRN_TRY(da.alloc(n));
    TRN_TRY(dout.alloc(n));
    TRN_TRY(upload(da.p, a, n, c->stream));
    if (b) {
        TRN_TRY(db.alloc(n));
        TRN_TRY(upload(db.p, b, n, c->stream));
    }
    TRN_TRY(launch_map(op, da.p, db.p, dout.p, n, c->stream));
    return download(out, dout.p, n, c->stream);
}

}  // namespace

extern "C" {

int trn_set_gemm_engine(int engine) {
    if (engine < 0 || engine > 3) return fail(TRN_INVALID_INPUT, "unknown GEMM engine %d", engine);
    g_engine.store(engine);
    return TRN_OK;
}
int trn_get_gemm_engine(void) { return g_engine.load(); }

// ================================ device-resident entry points ===================================
int trn_dot_f32_dev(const float* a, size_t na, const float* b, size_t nb, float* out, void* stream) {
    TRN_TRY(check_same_len(na, nb));
    TRN_TRY(need_ctx());
    return launch_reduce(Reduce::Dot, a, b, na, out, resolve_stream(stream));
}
int trn_sum_f32_dev(const float* a, size_t n, float* out, void* stream) {
    TRN_TRY(need_ctx());
    return launch_reduce(Reduce::Sum, a, nullptr, n, out, resolve_stream(stream));
}
int trn_sumsq_f32_dev(const float* a, size_t n, float* out, void* stream) {
    TRN_TRY(need_ctx());
    return launch_reduce(Reduce::SumSq, a, nullptr, n, out, resolve_stream(stream));
}
int trn_norm_l2_f32_dev(const float* a, size_t n, float* out, void* stream) {
    TRN_TRY(need_ctx());
    return launch_reduce(Reduce::NormL2, a, nullptr, n, out, resolve_stream(stream));
}
int trn_max_f32_dev(const float* a, size_t n, float* out, void* stream) {
    TRN_TRY(check_nonempty_invalid(n));
    TRN_TRY(need_ctx());
    return launch_argreduce(1, a, n, nullptr, out, resolve_stream(stream));
}
int trn_min_f32_dev(const float* a, size_t n, float* out, void* stream) {
    TRN_TRY(check_nonempty_invalid(n));
    TRN_TRY(need_ctx());
    return launch_argreduce(0, a, n, nullptr, out, resolve_stream(stream));
}
int trn_argmax_f32_dev(const float* a, size_t n, uint64_t* out, float* out_value, void* stream) {
    TRN_TRY(check_nonempty_invalid(n));
    TRN_TRY(need_ctx());
    return launch_argreduce(1, a, n, out, out_value, resolve_stream(stream));
}
int trn_argmin_f32_dev(const float* a, size_t n, uint64_t* out, float* out_value, void* stream) {
    TRN_TRY(check_nonempty_invalid(n));
    TRN_TRY(need_ctx());
    return launch_argreduce(0, a, n, out, out_value, resolve_stream(stream));
}
// Slice variants for sharded vectors: first_slice != 0 applies the a[0] seed rule,
// interior slices report "no candidate" as index UINT64_MAX / value = identity (-inf / +inf).
int trn_argmax_slice_f32_dev(const float* a, size_t n, int first_slice, uint64_t* out, float* out_value, void* stream) {
    TRN_TRY(check_nonempty_invalid(n));
    TRN_TRY(need_ctx());
    return launch_argreduce(1, a, n, out, out_value, resolve_stream(stream), first_slice != 0);
}
int trn_argmin_slice_f32_dev(const float* a, size_t n, int first_slice, uint64_t* out, float* out_value, void* stream) {
    TRN_TRY(check_nonempty_invalid(n));
    TRN_TRY(need_ctx());
    return launch_argreduce(0, a, n, out, out_value, resolve_stream(stream), first_slice != 0);
}
int trn_add_f32_dev(const float* a, size_t na, const float* b, size_t nb, float* out, void* stream) {
    TRN_TRY(check_same_len(na, nb));
    TRN_TRY(need_ctx());
    return launch_map(Map::Add, a, b, out, na, resolve_stream(stream));
}
int trn_mul_f32_dev(const float* a, size_t na, const float* b, size_t nb, float* out, void* stream) {
    TRN_TRY(check_same_len(na, nb));
    TRN_TRY(need_ctx());
    return launch_map(Map::Mul, a, b, out, na, resolve_stream(stream));
}
int trn_sigmoid_f32_dev(const float* a, size_t n, float* out, void* stream) {
    TRN_TRY(check_nonempty_emptyvec(n));
    TRN_TRY(need_ctx());
    return launch_map(Map::Sigmoid, a, nullptr, out, n, resolve_stream(stream));
}
int trn_gelu_f32_dev(const float* a, size_t n, float* out, void* stream) {
    TRN_TRY(check_nonempty_emptyvec(n));
    TRN_TRY(need_ctx());
    return launch_map(Map::Gelu, a, nullptr, out, n, resolve_stream(stream));
}
int trn_softmax_rows_f32_dev(const float* a, float* out, size_t rows, size_t cols, void* stream) {
    TRN_TRY(check_nonempty_emptyvec(rows * cols));
    TRN_TRY(need_ctx());
    return launch_softmax_rows(0, a, out, rows, cols, resolve_stream(stream));
}
int trn_log_softmax_rows_f32_dev(const float* a, float* out, size_t rows, size_t cols, void* stream) {
    TRN_TRY(check_nonempty_emptyvec(rows * cols));
    TRN_TRY(need_ctx());
    return launch_softmax_rows(1, a, out, rows, cols, resolve_stream(stream));
}
int trn_matmul_f32_dev(const float* a, size_t a_rows, size_t a_cols, const float* b, size_t b_rows, size_t b_cols,
                       float* c, void* stream) {
    TRN_TRY(check_matmul(a_rows, a_cols, b_rows, b_cols));
    TRN_TRY(need_ctx());
    cudaStream_t s = resolve_stream(stream);
    if (a_cols == 0 && a_rows * b_cols > 0) {
        TRN_CUDA(cudaMemsetAsync(c, 0, a_rows * b_cols * sizeof(float), s));
        return TRN_OK;
    }
    return gemm_dispatch(a, b, c, 1, a_rows, a_cols, b_cols, s);
}
int trn_batched_matmul_f32_dev(const float* a, size_t a_len, const float* b, size_t b_len, float* c, size_t batch,
                               size_t m, size_t k, size_t n, void* stream) {
    TRN_TRY(check_batched(a_len, b_len, batch, m, k, n));
    TRN_TRY(need_ctx());
    cudaStream_t s = resolve_stream(stream);
    if (k == 0 && batch * m * n > 0) {
        TRN_CUDA(cudaMemsetAsync(c, 0, batch * m * n * sizeof(float), s));
        return TRN_OK;
    }
    return gemm_dispatch(a, b, c, batch, m, k, n, s);
}
int trn_batched_matmul_4d_f32_dev(const float* a, size_t a_len, const float* b, size_t b_len, float* c,
                                  size_t batch, size_t heads, size_t m, size_t k, size_t n, void* stream) {
    TRN_TRY(check_batched_4d(a_len, b_len, batch, heads, m, k, n));
    TRN_TRY(need_ctx());
    cudaStream_t s = resolve_stream(stream);
    if (k == 0 && batch * heads * m * n > 0) {
        TRN_CUDA(cudaMemsetAsync(c, 0, batch * heads * m * n * sizeof(float), s));
        return TRN_OK;
    }
    return gemm_dispatch(a, b, c, batch * heads, m, k, n, s);
}
int trn_matvec_f32_dev(const float* a, size_t rows, size_t cols, const float* v, size_t v_len, float* y,
                       void* stream) {
    TRN_TRY(check_matvec(cols, v_len));
    TRN_TRY(need_ctx());
    cudaStream_t s = resolve_stream(stream);
    if (cols == 0 && rows > 0) {
        TRN_CUDA(cudaMemsetAsync(y, 0, rows * sizeof(float), s));
        return TRN_OK;
    }
    return launch_matvec(a, rows, cols, v, y, s);
}
int trn_transpose_f32_dev(const float* a, size_t rows, size_t cols, float* out, void* stream) {
    TRN_TRY(need_ctx());
    return launch_transpose(a, rows, cols, out, resolve_stream(stream));
}

// ==================================== host-slice entry points =====================================
int trn_dot_f32(const float* a, size_t na, const float* b, size_t nb, float* out) {
    TRN_TRY(check_same_len(na, nb));
    TRN_TRY(need_ctx());
    return host_reduce_f32(a, na, b, nb, out, [&](const float* da, const float* db, float* o, cudaStream_t s) {
        return launch_reduce(Reduce::Dot, da, db, na, o, s);
    });
}
int trn_sum_f32(const float* a, size_t n, float* out) {
    TRN_TRY(need_ctx());
    return host_reduce_f32(a, n, nullptr, 0, out, [&](const float* da, const float*, float* o, cudaStream_t s) {
        return launch_reduce(Reduce::Sum, da, nullptr, n, o, s);
    });
}
int trn_norm_l2_f32(const float* a, size_t n, float* out) {
    TRN_TRY(need_ctx());
    return host_reduce_f32(a, n, nullptr, 0, out, [&](const float* da, const float*, float* o, cudaStream_t s) {
        return launch_reduce(Reduce::NormL2, da, nullptr, n, o, s);
    });
}
int trn_max_f32(const float* a, size_t n, float* out) {
    TRN_TRY(check_nonempty_invalid(n));
    TRN_TRY(need_ctx());
    return host_arg(1, a, n, nullptr, out);
}
int trn_min_f32(const float* a, size_t n, float* out) {
    TRN_TRY(check_nonempty_invalid(n));
    TRN_TRY(need_ctx());
    return host_arg(0, a, n, nullptr, out);
}
int trn_argmax_f32(const float* a, size_t n, uint64_t* out) {
    TRN_TRY(check_nonempty_invalid(n));
    TRN_TRY(need_ctx());
    return host_arg(1, a, n, out, nullptr);
}
int trn_argmin_f32(const float* a, size_t n, uint64_t* out) {
    TRN_TRY(check_nonempty_invalid(n));
    TRN_TRY(need_ctx());
    return host_arg(0, a, n, out, nullptr);
}
int trn_add_f32(const float* a, size_t na, const float* b, size_t nb, float* out) {
    TRN_TRY(check_same_len(na, nb));
    TRN_TRY(need_ctx());
    return host_map(Map::Add, a, b, out, na);
}
int trn_mul_f32(const float* a, size_t na, const float* b, size_t nb, float* out) {
    TRN_TRY(check_same_len(na, nb));
    TRN_TRY(need_ctx());
    return host_map(Map::Mul, a, b, out, na);
}
int trn_sigmoid_f32(const float* a, size_t n, float* out) {
    TRN_TRY(check_nonempty_emptyvec(n));
    TRN_TRY(need_ctx());
    return host_map(Map::Sigmoid, a, nullptr, out, n);
}
int trn_gelu_f32(const float* a, size_t n, float* out) {
    TRN_TRY(check_nonempty_emptyvec(n));
    TRN_TRY(need_ctx());
    return host_map(Map::Gelu, a, nullptr, out, n);
}

static int host_softmax(int log_variant, const float* a, float* out, size_t rows, size_t cols) {
    TRN_TRY(check_nonempty_emptyvec(rows * cols));
    TRN_TRY(need_ctx());
    Context* c = ctx();
    const size_t n = rows * cols;
    DevTemp da(c->stream), dout(c->stream);
    TRN_TRY(da.alloc(n));
    TRN_TRY(dout.alloc(n));
    TRN_TRY(upload(da.p, a, n, c->stream));
    TRN_TRY(launch_softmax_rows(log_variant, da.p, dout.p, rows, cols, c->stream));
    return download(out, dout.p, n, c->stream);
}
int trn_softmax_rows_f32(const float* a, float* out, size_t rows, size_t cols) {
    return host_softmax(0, a, out, rows, cols);
}
int trn_log_softmax_rows_f32(const float* a, float* out, size_t rows, size_t cols) {
    return host_softmax(1, a, out, rows, cols);
}

static int host_gemm(const float* a, const float* b, float* c, size_t batch, size_t m, size_t k, size_t n) {
    Context* cx = ctx();
    const size_t na = batch * m * k, nb = batch * k * n, nc = batch * m * n;
    if (nc == 0) return TRN_OK;
    DevTemp da(cx->stream), db(cx->stream), dc(cx->stream);
    TRN_TRY(da.alloc(na));
    TRN_TRY(db.alloc(nb));
    TRN_TRY(dc.alloc(nc));
    TRN_TRY(upload(da.p, a, na, cx->stream));
    TRN_TRY(upload(db.p, b, nb, cx->stream));
    if (k == 0) TRN_CUDA(cudaMemsetAsync(dc.p, 0, nc * sizeof(float), cx->stream));
    else TRN_TRY(gemm_dispatch(da.p, db.p, dc.p, batch, m, k, n, cx->stream));
    return download(c, dc.p, nc, cx->stream);
}
int trn_matmul_f32(const float* a, size_t a_rows, size_t a_cols, const float* b, size_t b_rows, size_t b_cols,
                   float* c) {
    TRN_TRY(check_matmul(a_rows, a_cols, b_rows, b_cols));
    TRN_TRY(need_ctx());
    return host_gemm(a, b, c, 1, a_rows, a_cols, b_cols);
}
int trn_batched_matmul_f32(const float* a, size_t a_len, const float* b, size_t b_len, float* c, size_t batch,
                           size_t m, size_t k, size_t n) {
    TRN_TRY(check_batched(a_len, b_len, batch, m, k, n));
    TRN_TRY(need_ctx());
    return host_gemm(a, b, c, batch, m, k, n);
}
int trn_batched_matmul_4d_f32(const float* a, size_t a_len, const float* b, size_t b_len, float* c, size_t batch,
                              size_t heads, size_t m, size_t k, size_t n) {
    TRN_TRY(check_batched_4d(a_len, b_len, batch, heads, m, k, n));
    TRN_TRY(need_ctx());
    return host_gemm(a, b, c, batch * heads, m, k, n);
}
int trn_matvec_f32(const float* a, size_t rows, size_t cols, const float* v, size_t v_len, float* y) {
    TRN_TRY(check_matvec(cols, v_len));
    TRN_TRY(need_ctx());
    Context* c = ctx();
    if (rows == 0) return TRN_OK;
    DevTemp da(c->stream), dv(c->stream), dy(c->stream);
    TRN_TRY(da.alloc(rows * cols));
    TRN_TRY(dv.alloc(cols));
    TRN_TRY(dy.alloc(rows));
    TRN_TRY(upload(da.p, a, rows * cols, c->stream));
    TRN_TRY(upload(dv.p, v, cols, c->stream));
    if (cols == 0) TRN_CUDA(cudaMemsetAsync(dy.p, 0, rows * sizeof(float), c->stream));
    else TRN_TRY(launch_matvec(da.p, rows, cols, dv.p, dy.p, c->stream));
    return download(y, dy.p, rows, c->stream);
}
int trn_transpose_f32(const float* a, size_t rows, size_t cols, float* out) {
    TRN_TRY(need_ctx());
    Context* c = ctx();
    const size_t n = rows * cols;
    if (n == 0) return TRN_OK;
    DevTemp da(c->stream), dout(c->stream);
    TRN_TRY(da.alloc(n));
    TRN_TRY(dout.alloc(n));
    TRN_TRY(upload(da.p, a, n, c->stream));
    TRN_TRY(launch_transpose(da.p, rows, cols, dout.p, c->stream));
    return download(out, dout.p, n, c->stream);
}

}  // extern "C"
