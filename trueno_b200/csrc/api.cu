// api.cu — the extern "C" operator surface (include/trueno_cuda.h): validation with the
// reference's exact error values/messages, engine dispatch, and the host-slice wrappers that
// stage through HBM.  No entry point ever computes on the CPU.
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"

using namespace trn;

// pre-split right-hand operand (trn_gemm_prepare_b_dev, below)
struct trn_gemm_b {
    const float* b = nullptr;   // borrowed: the caller's B in HBM
    size_t k = 0, n = 0;
    float* split = nullptr;     // B_hi then B_lo, [n][kpad] each; null when no kernel would read them
    int* flag = nullptr;        // raised by the split when B holds an Inf / NaN
};

namespace {

// Rust's `{}` for f32: shortest representation that round-trips, integers without a fraction ("1", "0.5", "-2.25")
std::string fmt_f32(float v) {
    if (v != v) return "NaN";                       // Rust's Display for f32
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[64];
    for (int prec = 1; prec <= 9; ++prec) {
        snprintf(buf, sizeof buf, "%.*g", prec, (double)v);
        if (strtof(buf, nullptr) == v) break;
    }
    std::string r(buf);
    if (r.find('e') != std::string::npos) {   // Rust never prints exponents for f32 Display
        snprintf(buf, sizeof buf, "%.60f", (double)v);   // Rust prints the shortest round-tripping digits, positionally
        r = buf;
        // cut to the shortest prefix that still round-trips
        const size_t dot = r.find('.');
        for (size_t len = dot + 2; len <= r.size(); ++len)
            if (strtof(r.substr(0, len).c_str(), nullptr) == v) { r = r.substr(0, len); break; }
        while (r.find('.') != std::string::npos && (r.back() == '0' || r.back() == '.')) { const bool dot = r.back() == '.'; r.pop_back(); if (dot) break; }
    }
    return r;
}

std::atomic<int> g_engine{0};

// Host-slice calls stage through the ONE backend stream and its per-stream scalar slots / pinned staging ring,
// so concurrent callers (the reference's rayon workers may call a backend from several threads) are serialised
// here.  `_dev` calls take caller-owned memory and streams and need no lock.
#define TRN_HOST_LOCK() std::lock_guard<std::mutex> _host_lock(trn::host_mutex())

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

}  // namespace

// auto: the tensor-core tile is 128x256; below that the SIMT kernel wins.  The choice depends on
// (m, k, n) only — never on batch or on how the host path blocks the rows — so a batched product is
// bit-identical to the loop of single products the reference runs (src/matrix.rs:507-524).
bool trn::gemm_auto_uses_tc(size_t m, size_t k, size_t n) {
    return gemm_tc_supported(m, k, n) && m >= 128 && n >= 128 && k >= 32 && m * n * k >= (size_t)1 << 24;
}

namespace {

// ---- GEMM engine dispatch -----------------------------------------------------------------------
// Matrix::matmul routes by shape (src/matrix.rs:293-356); the CUDA backend routes every shape to
// the device (north star: unconditional, no thresholds that fall back to the CPU) and only picks
// WHICH device kernel: rows == 1 -> vecmat (reference quirk preserved); tiles that fill a tcgen05
// tile -> 3xTF32; the rest -> SIMT FFMA.
// route_m != 0: `a` holds a row block of a product with route_m rows, and the kernel is picked as for that whole product.
int gemm_dispatch(const float* a, const float* b, float* c, size_t batch, size_t m, size_t k, size_t n,
                  cudaStream_t s, size_t route_m = 0) {
    if (batch == 0 || m == 0 || n == 0) return TRN_OK;
    const size_t mr = route_m ? route_m : m;
    if (mr == 1 && k > 0) {
        for (size_t i = 0; i < batch; ++i) TRN_TRY(launch_vecmat(a + i * k, b + i * k * n, k, n, c + i * n, s));
        return TRN_OK;
    }
    const int engine = g_engine.load();
    if (engine == 1) return launch_gemm_simt(a, b, c, batch, m, k, n, s);
    if (engine == 2 || engine == 3) {
        if (!gemm_tc_supported(mr, k, n))
            return fail(TRN_INVALID_INPUT, "tcgen05 GEMM engine forced for an unsupported shape %zux%zux%zu", mr, k, n);
        return launch_gemm_tc(a, b, c, batch, m, k, n, engine == 2 ? 3 : 1, s, route_m);
    }
    if (gemm_auto_uses_tc(mr, k, n)) return launch_gemm_tc(a, b, c, batch, m, k, n, 3, s, route_m);
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
    TRN_HOST_LOCK();
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
    TRN_HOST_LOCK();
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

// ---- pipelined host-slice maps and row kernels ---------------------------------------------------------------------------
// An element-wise op on host slices is two PCIe transfers around a kernel that is ~50x shorter than either of them.  From
// pinned memory the three legs run as a pipeline over three streams, chunk by chunk — H2D of chunk i+1, kernel on chunk i,
// D2H of chunk i-1 — so the link works in BOTH directions at once (131 M-element gelu: 19.5 -> ~10.5 ms).  `launch(off,
// cnt, stream)` enqueues the kernel for elements [off, off + cnt) of the device copies; the op is element- or row-wise, so
// the result is bit-identical to the unchunked call.
struct CudaEvent {
    cudaEvent_t e = nullptr;
    ~CudaEvent() { if (e) cudaEventDestroy(e); }
};

template <class F>
int host_pipeline(const float* const* inputs, int n_in, float* out, size_t n, size_t chunk, F&& launch, float* const* dev_in, float* dev_out) {
    Context* c = ctx();
    cudaStream_t s_main = c->stream, s_up = c->copy_stream, s_down = c->d2h_stream;
    const size_t nchunks = (n + chunk - 1) / chunk;
    std::vector<CudaEvent> up(nchunks), done(nchunks);
    CudaEvent ready;
    auto enqueue = [&]() -> int {
        for (auto& e : up) TRN_CUDA(cudaEventCreateWithFlags(&e.e, cudaEventDisableTiming));
        for (auto& e : done) TRN_CUDA(cudaEventCreateWithFlags(&e.e, cudaEventDisableTiming));
        TRN_CUDA(cudaEventCreateWithFlags(&ready.e, cudaEventDisableTiming));
        TRN_CUDA(cudaEventRecord(ready.e, s_main));            // the stream-ordered allocations are complete
        TRN_CUDA(cudaStreamWaitEvent(s_up, ready.e, 0));
        TRN_CUDA(cudaStreamWaitEvent(s_down, ready.e, 0));
        for (size_t i = 0; i < nchunks; ++i) {
            const size_t off = i * chunk, cnt = n - off < chunk ? n - off : chunk;
            for (int k = 0; k < n_in; ++k)
                TRN_CUDA(cudaMemcpyAsync(dev_in[k] + off, inputs[k] + off, cnt * sizeof(float), cudaMemcpyHostToDevice, s_up));
            TRN_CUDA(cudaEventRecord(up[i].e, s_up));
            TRN_CUDA(cudaStreamWaitEvent(s_main, up[i].e, 0));
            TRN_TRY(launch(off, cnt, s_main));
            TRN_CUDA(cudaEventRecord(done[i].e, s_main));
            TRN_CUDA(cudaStreamWaitEvent(s_down, done[i].e, 0));
            TRN_CUDA(cudaMemcpyAsync(out + off, dev_out + off, cnt * sizeof(float), cudaMemcpyDeviceToHost, s_down));
        }
        return TRN_OK;
    };
    const int st = enqueue();
    // the host slices are borrowed: whatever happened, nothing may still be in flight when the call returns
    const cudaError_t e1 = cudaStreamSynchronize(s_up), e2 = cudaStreamSynchronize(s_main), e3 = cudaStreamSynchronize(s_down);
    if (st != TRN_OK) return st;
    TRN_CUDA(e1);
    TRN_CUDA(e2);
    TRN_CUDA(e3);
    return TRN_OK;
}
constexpr size_t kPipeChunk = (size_t)4 << 20;   // elements per chunk: 16 MiB per operand (~0.3 ms on the link)

int host_map(Map op, const float* a, const float* b, const float* c3, float* out, size_t n, float p0 = 0.f, float p1 = 0.f) {
    TRN_HOST_LOCK();
    Context* c = ctx();
    if (n == 0) return TRN_OK;
    if (n >= 2 * kPipeChunk && is_pinned_host(a) && (!b || is_pinned_host(b)) && (!c3 || is_pinned_host(c3)) && is_pinned_host(out)) {
        DevTemp da(c->stream), db(c->stream), dc(c->stream), dout(c->stream);
        TRN_TRY(da.alloc(n));
        TRN_TRY(dout.alloc(n));
        if (b) TRN_TRY(db.alloc(n));
        if (c3) TRN_TRY(dc.alloc(n));
        const float* ins[3] = {a, b, c3};
        float* devs[3] = {da.p, db.p, dc.p};
        int n_in = 1;
        if (b) { ins[n_in] = b; devs[n_in] = db.p; ++n_in; }
        if (c3) { ins[n_in] = c3; devs[n_in] = dc.p; ++n_in; }
        return host_pipeline(ins, n_in, out, n, kPipeChunk, [&](size_t off, size_t cnt, cudaStream_t s) {
            return launch_map(op, da.p + off, b ? db.p + off : nullptr, c3 ? dc.p + off : nullptr, dout.p + off, cnt, p0, p1, s);
        }, devs, dout.p);
    }
    DevTemp da(c->stream), db(c->stream), dc(c->stream), dout(c->stream);
    TRN_TRY(da.alloc(n));
    TRN_TRY(dout.alloc(n));
    TRN_TRY(upload(da.p, a, n, c->stream));
    if (b) {
        TRN_TRY(db.alloc(n));
        TRN_TRY(upload(db.p, b, n, c->stream));
    }
    if (c3) {
        TRN_TRY(dc.alloc(n));
        TRN_TRY(upload(dc.p, c3, n, c->stream));
    }
    TRN_TRY(launch_map(op, da.p, db.p, dc.p, dout.p, n, p0, p1, c->stream));
    return download(out, dout.p, n, c->stream);
}

int check_clamp(float lo, float hi) {  // src/vector.rs:2964-2969
    if (lo > hi) return fail(TRN_INVALID_INPUT, "Invalid clamp range: min (%s) > max (%s)", fmt_f32(lo).c_str(), fmt_f32(hi).c_str());
    return TRN_OK;
}

}  // namespace

extern "C" {

static int host_gemm_pipelined(const float* a, const float* b, float* c, size_t batch, size_t m, size_t k, size_t n,
                               const trn_gemm_b* prepared);

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
int trn_argmax_slice_pair_f32_dev(const float* a, size_t n, uint64_t slice_start, trn_arg_pair* out, void* stream) {
    if (slice_start == 0) TRN_TRY(check_nonempty_invalid(n));   // an empty INTERIOR slice reports "no candidate"
    TRN_TRY(need_ctx());
    return launch_argreduce(1, a, n, nullptr, nullptr, resolve_stream(stream), slice_start == 0, slice_start, out);
}
int trn_argmin_slice_pair_f32_dev(const float* a, size_t n, uint64_t slice_start, trn_arg_pair* out, void* stream) {
    if (slice_start == 0) TRN_TRY(check_nonempty_invalid(n));   // an empty INTERIOR slice reports "no candidate"
    TRN_TRY(need_ctx());
    return launch_argreduce(0, a, n, nullptr, nullptr, resolve_stream(stream), slice_start == 0, slice_start, out);
}
int trn_arg_combine_f32_dev(const trn_arg_pair* pairs, size_t count, int is_max, uint64_t* out_idx, float* out_value,
                            void* stream) {
    TRN_TRY(need_ctx());
    return launch_arg_combine(pairs, count, is_max, out_idx, out_value, resolve_stream(stream));
}
int trn_add_f32_dev(const float* a, size_t na, const float* b, size_t nb, float* out, void* stream) {
    TRN_TRY(check_same_len(na, nb));
    TRN_TRY(need_ctx());
    return launch_map(Map::Add, a, b, nullptr, out, na, 0.f, 0.f, resolve_stream(stream));
}
int trn_mul_f32_dev(const float* a, size_t na, const float* b, size_t nb, float* out, void* stream) {
    TRN_TRY(check_same_len(na, nb));
    TRN_TRY(need_ctx());
    return launch_map(Map::Mul, a, b, nullptr, out, na, 0.f, 0.f, resolve_stream(stream));
}
int trn_sigmoid_f32_dev(const float* a, size_t n, float* out, void* stream) {
    TRN_TRY(check_nonempty_emptyvec(n));
    TRN_TRY(need_ctx());
    return launch_map(Map::Sigmoid, a, nullptr, nullptr, out, n, 0.f, 0.f, resolve_stream(stream));
}
int trn_gelu_f32_dev(const float* a, size_t n, float* out, void* stream) {
    TRN_TRY(check_nonempty_emptyvec(n));
    TRN_TRY(need_ctx());
    return launch_map(Map::Gelu, a, nullptr, nullptr, out, n, 0.f, 0.f, resolve_stream(stream));
}
int trn_softmax_rows_f32_dev(const float* a, float* out, size_t rows, size_t cols, void* stream) {
    TRN_TRY(check_nonempty_emptyvec(rows * cols));
    TRN_TRY(need_ctx());
    return launch_softmax_rows(0, a, out, rows, cols, resolve_stream(stream));
}
// One Vector::softmax / log_softmax sharded over several GPUs: stats of this rank's slice, then — after
// the ranks have all_gathered their pairs — normalisation of the slice by the fold of all pairs.
int trn_softmax_slice_stats_f32_dev(const float* a, size_t n, float* pair_out, void* stream) {
    TRN_TRY(need_ctx());
    if (n == 0) {   // an empty slice of a sharded vector takes part with the identity pair (-inf, 0)
        static const float kIdentity[2] = {-INFINITY, 0.0f};
        TRN_CUDA(cudaMemcpyAsync(pair_out, kIdentity, sizeof kIdentity, cudaMemcpyHostToDevice, resolve_stream(stream)));
        return TRN_OK;
    }
    return launch_softmax_slice_stats(a, n, pair_out, resolve_stream(stream));
}
int trn_softmax_slice_apply_f32_dev(const float* a, size_t n, const float* pairs, size_t npairs, int log_variant, float* out,
                                    void* stream) {
    if (n == 0) return TRN_OK;   // nothing to write for an empty slice
    if (npairs == 0) return fail(TRN_INVALID_INPUT, "no (max, sum) pairs to normalise by");
    TRN_TRY(need_ctx());
    return launch_softmax_slice_apply(a, n, pairs, npairs, log_variant, out, resolve_stream(stream));
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
// ---- pre-split right-hand operand (repeated products against the same B) ------------------------------------------
// Matrix::matmul re-reads `other` on every call (the CPU path even re-transposes it, src/matrix.rs:934); on the device the
// tensor-core path re-SPLITS it: 2 x |B| of writes per call (2.4 ms for the 4 GiB B of a 32768^2 row-block shard).  A
// handle keeps B_hi / B_lo (K-major, padded) in HBM, so every later product pays for A only.  Shapes that take another
// kernel (fused-split, SIMT, vecmat) need nothing prepared: the handle then just remembers B.  B itself is BORROWED: it
// must stay alive and unchanged while the handle is used (the IEEE fallback for Inf/NaN inputs and the non-tensor kernels
// read it).  Results are bit-identical to trn_matmul_f32_dev(a, b).
int trn_gemm_prepare_b_dev(const float* b, size_t b_rows, size_t b_cols, trn_gemm_b** out, void* stream) {
    if (!out) return fail(TRN_INVALID_INPUT, "trn_gemm_prepare_b_dev: null output handle");
    TRN_TRY(need_ctx());
    cudaStream_t s = resolve_stream(stream);
    trn_gemm_b* h = new trn_gemm_b();
    h->b = b;
    h->k = b_rows;
    h->n = b_cols;
    const size_t kpad = gemm_tc_kpad(b_rows);
    const size_t elems = b_cols * kpad;
    // prepared only for the pre-pass tensor-core kernel: gemm_prepared_uses_split() repeats this test per product
    if (b_rows > 0 && b_cols > 0 && gemm_tc_supported(256, b_rows, b_cols) && !gemm_tc_uses_fused(nullptr, b, 256, b_rows, b_cols)) {
        cudaError_t e = cudaMalloc(&h->split, (2 * elems) * sizeof(float) + 256);
        if (e != cudaSuccess) { delete h; return fail_cuda(e, "cudaMalloc (pre-split B)"); }
        h->flag = reinterpret_cast<int*>(h->split + 2 * elems);
        int st = cudaMemsetAsync(h->flag, 0, sizeof(int), s) == cudaSuccess ? TRN_OK : fail(TRN_GPU_ERROR, "cudaMemsetAsync failed");
        if (st == TRN_OK) st = gemm_tc_split_b(b, h->split, h->split + elems, 1, b_rows, b_cols, h->flag, s);
        if (st != TRN_OK) { cudaFree(h->split); delete h; return st; }
    }
    *out = h;
    return TRN_OK;
}

int trn_gemm_b_free(trn_gemm_b* h) {
    if (!h) return TRN_OK;
    cudaError_t e = h->split ? cudaFree(h->split) : cudaSuccess;   // cudaFree waits for the work that still reads it
    delete h;
    if (e != cudaSuccess) return fail_cuda(e, "cudaFree (pre-split B)");
    return TRN_OK;
}

static int matmul_prepared_dev(const float* a, size_t a_rows, size_t a_cols, const trn_gemm_b* bh, float* c, void* stream,
                               size_t route_m, const char* who) {
    if (!bh) return fail(TRN_INVALID_INPUT, "%s: null B handle", who);
    TRN_TRY(check_matmul(a_rows, a_cols, bh->k, bh->n));
    TRN_TRY(need_ctx());
    cudaStream_t s = resolve_stream(stream);
    const size_t m = a_rows, k = a_cols, n = bh->n;
    const size_t mr = route_m ? route_m : m;
    if (k == 0 && m * n > 0) {
        TRN_CUDA(cudaMemsetAsync(c, 0, m * n * sizeof(float), s));
        return TRN_OK;
    }
    // the same routing as trn_matmul_f32_dev; only the pre-pass tensor-core product has something to reuse
    const int engine = g_engine.load();
    const bool tc3 = mr > 1 && (engine == 2 || (engine == 0 && gemm_auto_uses_tc(mr, k, n)));
    if (!bh->split || !tc3 || gemm_tc_uses_fused(a, bh->b, mr, k, n)) return gemm_dispatch(a, bh->b, c, 1, m, k, n, s, route_m);
    const size_t kpad = gemm_tc_kpad(k);
    float* scratch = nullptr;
    TRN_TRY(scratch_alloc((void**)&scratch, 2 * m * kpad * sizeof(float) + 256, s));
    int* flag = reinterpret_cast<int*>(scratch + 2 * m * kpad);
    // start from B's verdict, then let the split of A add its own
    int st = cudaMemcpyAsync(flag, bh->flag, sizeof(int), cudaMemcpyDeviceToDevice, s) == cudaSuccess ? TRN_OK : fail(TRN_GPU_ERROR, "cudaMemcpyAsync failed");
    gemm_profile_begin(s);
    if (st == TRN_OK) st = gemm_tc_split_a(a, scratch, scratch + m * kpad, 1, m, k, flag, s);
    gemm_profile_mid(s);
    if (st == TRN_OK) st = gemm_tc_main(scratch, scratch + m * kpad, bh->split, bh->split + n * kpad, c, 1, m, k, n, 3, flag, s, route_m);
    gemm_profile_end(s);
    if (st == TRN_OK) st = launch_gemm_simt(a, bh->b, c, 1, m, k, n, s, flag);
    scratch_free(scratch, s);
    return st;
}
int trn_matmul_prepared_f32_dev(const float* a, size_t a_rows, size_t a_cols, const trn_gemm_b* bh, float* c, void* stream) {
    return matmul_prepared_dev(a, a_rows, a_cols, bh, c, stream, 0, "trn_matmul_prepared_f32_dev");
}

// ---- row blocks of one product (Matrix::matmul sharded by C row blocks, src/matrix.rs:962-1011) ----------------
// `a` holds block_rows consecutive rows of an A with total_rows rows.  The kernel family is chosen as trn_matmul_f32_dev
// chooses it for the WHOLE product, whatever the block's own height (44 rows of a 1324-row product stay on the tensor
// cores, one row of it does not become the vecmat special case): every kernel computes a row of C with the same
// arithmetic in whichever tile the row lies, so the gathered row blocks are bit-identical to the unsharded product.
static int check_rowblock(size_t block_rows, size_t total_rows) {
    if (block_rows > total_rows)
        return fail(TRN_INVALID_INPUT, "row block of %zu rows exceeds the %zu rows of the matrix", block_rows, total_rows);
    return TRN_OK;
}
int trn_matmul_rowblock_f32_dev(const float* a, size_t block_rows, size_t total_rows, size_t a_cols, const float* b, size_t b_rows,
                                size_t b_cols, float* c, void* stream) {
    TRN_TRY(check_matmul(total_rows, a_cols, b_rows, b_cols));
    TRN_TRY(check_rowblock(block_rows, total_rows));
    TRN_TRY(need_ctx());
    cudaStream_t s = resolve_stream(stream);
    if (a_cols == 0 && block_rows * b_cols > 0) {
        TRN_CUDA(cudaMemsetAsync(c, 0, block_rows * b_cols * sizeof(float), s));
        return TRN_OK;
    }
    return gemm_dispatch(a, b, c, 1, block_rows, a_cols, b_cols, s, total_rows);
}
int trn_matmul_rowblock_prepared_f32_dev(const float* a, size_t block_rows, size_t total_rows, size_t a_cols, const trn_gemm_b* bh,
                                         float* c, void* stream) {
    TRN_TRY(check_rowblock(block_rows, total_rows));
    if (block_rows == 0) return TRN_OK;
    return matmul_prepared_dev(a, block_rows, a_cols, bh, c, stream, total_rows, "trn_matmul_rowblock_prepared_f32_dev");
}

// host-slice twin: A and C are host slices, B stays resident.  Pinned slices of large tensor-core products are pipelined
// (row blocks of A up, GEMM, row blocks of C down, on three streams); results are bit-identical to the resident call.
int trn_matmul_prepared_f32(const float* a, size_t a_rows, size_t a_cols, const trn_gemm_b* bh, float* c) {
    if (!bh) return fail(TRN_INVALID_INPUT, "trn_matmul_prepared_f32: null B handle");
    TRN_TRY(check_matmul(a_rows, a_cols, bh->k, bh->n));
    TRN_TRY(need_ctx());
    TRN_HOST_LOCK();
    Context* cx = ctx();
    const size_t m = a_rows, k = a_cols, n = bh->n;
    if (m * n == 0) return TRN_OK;
    if (g_engine.load() == 0 && k > 0 && m >= 1024 && gemm_auto_uses_tc(m, k, n) && (m * k + m * n) * sizeof(float) >= ((size_t)64 << 20) &&
        (bh->split || gemm_tc_uses_fused(nullptr, bh->b, m, k, n)) && is_pinned_host(a) && is_pinned_host(c))
        return host_gemm_pipelined(a, nullptr, c, 1, m, k, n, bh);
    DevTemp da(cx->stream), dc(cx->stream);
    TRN_TRY(da.alloc(m * k));
    TRN_TRY(dc.alloc(m * n));
    TRN_TRY(upload(da.p, a, m * k, cx->stream));
    TRN_TRY(trn_matmul_prepared_f32_dev(da.p, m, k, bh, dc.p, cx->stream));
    return download(c, dc.p, m * n, cx->stream);
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
    return host_map(Map::Add, a, b, nullptr, out, na);
}
int trn_mul_f32(const float* a, size_t na, const float* b, size_t nb, float* out) {
    TRN_TRY(check_same_len(na, nb));
    TRN_TRY(need_ctx());
    return host_map(Map::Mul, a, b, nullptr, out, na);
}
int trn_sigmoid_f32(const float* a, size_t n, float* out) {
    TRN_TRY(check_nonempty_emptyvec(n));
    TRN_TRY(need_ctx());
    return host_map(Map::Sigmoid, a, nullptr, nullptr, out, n);
}
int trn_gelu_f32(const float* a, size_t n, float* out) {
    TRN_TRY(check_nonempty_emptyvec(n));
    TRN_TRY(need_ctx());
    return host_map(Map::Gelu, a, nullptr, nullptr, out, n);
}

static int host_softmax(int log_variant, const float* a, float* out, size_t rows, size_t cols) {
    TRN_TRY(check_nonempty_emptyvec(rows * cols));
    TRN_TRY(need_ctx());
    TRN_HOST_LOCK();
    Context* c = ctx();
    const size_t n = rows * cols;
    DevTemp da(c->stream), dout(c->stream);
    TRN_TRY(da.alloc(n));
    TRN_TRY(dout.alloc(n));
    // many rows from pinned memory: pipeline whole-row chunks (the row kernels for cols <= 32768 are chosen by the row
    // length alone, so chunking the rows changes no bit)
    const size_t rows_per_chunk = cols ? kPipeChunk / cols : 0;
    if (cols <= 32768 && cols % 4 == 0 && rows_per_chunk >= 16 && rows >= 2 * rows_per_chunk && is_pinned_host(a) && is_pinned_host(out)) {
        const float* ins[1] = {a};
        float* devs[1] = {da.p};
        return host_pipeline(ins, 1, out, n, rows_per_chunk * cols, [&](size_t off, size_t cnt, cudaStream_t s) {
            return launch_softmax_rows(log_variant, da.p + off, dout.p + off, cnt / cols, cols, s);
        }, devs, dout.p);
    }
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

// ---- pipelined host-slice GEMM -----------------------------------------------------------------------
// The host-slice call is bound by PCIe, not by the tensor cores (8192^3: 768 MiB over the link vs
// ~4.5 ms of math), so the transfers are overlapped with the math instead of bracketing it:
//   copy stream : H2D of B, then of A in row blocks (single product) / of A and B per group of heads (batched)
//   main stream : split B once (or per group), then per block: split A block -> tcgen05 GEMM -> (IEEE fallback)
//   d2h stream  : D2H of each finished C block while the next block is uploaded and computed
// H2D and D2H run on different copy engines in opposite directions of the link.  Every block runs the same
// kernels with the same k-order as the unblocked call: results are bit-identical to trn_matmul_f32_dev.
static bool pipe_trace() {   // TRN_PIPE_TRACE=1: timing events in the pipelined GEMM, phase times on stderr (experiments)
    static const bool on = [] { const char* e = getenv("TRN_PIPE_TRACE"); return e && atoi(e) != 0; }();
    return on;
}
struct Event {
    cudaEvent_t e = nullptr;
    int create() { return cudaEventCreateWithFlags(&e, pipe_trace() ? cudaEventDefault : cudaEventDisableTiming) == cudaSuccess ? TRN_OK : fail(TRN_GPU_ERROR, "cudaEventCreate failed"); }
    ~Event() { if (e) cudaEventDestroy(e); }
};

// `prepared` != nullptr (single products only): B is already resident — and split, when the kernel wants that — so only
// the row blocks of A go up and the row blocks of C come down.
static int host_gemm_pipelined(const float* a, const float* b, float* c, size_t batch, size_t m, size_t k, size_t n,
                               const trn_gemm_b* prepared) {
    Context* cx = ctx();
    cudaStream_t s_main = cx->stream, s_up = cx->copy_stream, s_down = cx->d2h_stream;
    // (Alternating the row blocks over TWO compute streams, so that a block's 22-tile second wave does not leave 52 CTA pairs
    // idle behind a kernel boundary, measured SLOWER: 11.72 vs 11.40 ms at 8192^3 — the concurrent split / GEMM kernels of
    // neighbouring blocks slow each other and the uploads.  TRN_PIPE_TRACE=1 prints the phase times of a call.)
    const size_t kpad = gemm_tc_kpad(k);
    const bool single = batch == 1;
    // the same kernel choice as the resident call (gemm_tc.cu): results stay bit-identical to it.  Scratch sub-buffers
    // are 256-byte aligned, so alignment never differs from an aligned resident operand.
    const bool fused = gemm_tc_uses_fused(nullptr, nullptr, m, k, n);
    // work units: row blocks of one product, or groups of whole (batch*head) products
    size_t units = batch, per_block = 1;
    std::vector<size_t> row_start;   // single product: start row of every block, plus m as the sentinel
    if (single) {
        // Row blocks of whole 256-row pair tiles.  After B has arrived the three stages of a block — upload of its A
        // rows, split + GEMM, download of its C rows — run as a pipeline over three streams; the end-to-end time is
        // B upload + (blocks + 2) x the slowest stage, so blocks should be as small as the GEMM stays efficient.
        // Measured on 8192^3 (TRN_PIPE_ROWS), round 1: 256-row blocks 13.0 ms, 512 11.4, 768 11.15, 1024 11.5, 2304 12.6;
        // round 2 (with the phase trace, same box): 512 11.09, 768 11.37, 1024 11.87.  The PCIe floor (512 MiB up + 256 MiB
        // down concurrently) is 10.3 ms.
        static const size_t forced_rows = [] { const char* e = getenv("TRN_PIPE_ROWS"); return e ? (size_t)atol(e) : (size_t)0; }();
        // (One-wave blocks — 512 rows at n = 8192 are 64 CTA-pair tiles where 768 rows are 96 = 1.3 waves that take two — look
        // better in a single phase trace, 11.09 vs 11.37 ms, but are bimodal over repeated processes on one box: 11.2-12.3 ms
        // against a steady 11.4-11.6 for 768-row blocks.  Eleven blocks stay.)
        size_t unit_rows = forced_rows ? (forced_rows + 255) / 256 * 256 : 256 * ((m / 11 + 255) / 256);
        if (unit_rows < 256) unit_rows = 256;
        size_t r = 0;
        while (m - r > unit_rows) { row_start.push_back(r); r += unit_rows; }
        // the tail: [rest, 256, 256] rows — after the last upload only a 256-row GEMM and its 8 MiB D2H remain
        const size_t last = m - r;
        row_start.push_back(r);
        if (last >= 768) {
            const size_t rest = (last - 512 + 255) / 256 * 256;
            row_start.push_back(r + rest);
            if (last - rest > 256) row_start.push_back(r + rest + 256);
        }
        // the fused-split kernel needs more than one 128-row tile per block: a short tail joins the block before it
        while (fused && row_start.size() > 1 && m - row_start.back() <= 128) row_start.pop_back();
        units = row_start.size();
        row_start.push_back(m);
    } else {
        const size_t bytes_per = (m * k + k * n + m * n) * sizeof(float);
        per_block = ((size_t)64 << 20) / (bytes_per ? bytes_per : 1);   // ~64 MiB of traffic per group
        if (per_block < 1) per_block = 1;
        if (per_block > batch) per_block = batch;
        units = (batch + per_block - 1) / per_block;
    }

    const size_t na = batch * m * k, nb = prepared ? 0 : batch * k * n, nc = batch * m * n;
    const size_t a_split = fused ? 0 : batch * m * kpad, b_split = (fused || prepared) ? 0 : batch * n * kpad;
    // every sub-buffer starts on a 256-byte boundary (TMA needs 16-byte aligned tensor bases)
    auto pad64 = [](size_t x) { return (x + 63) & ~(size_t)63; };
    float* dev = nullptr;
    TRN_TRY(scratch_alloc((void**)&dev, (pad64(na) + pad64(nb) + pad64(nc) + 2 * pad64(a_split) + 2 * pad64(b_split)) * sizeof(float) + 256, s_main));
    float* da = dev;
    float* db = da + pad64(na);
    float* dc = db + pad64(nb);
    float* a_hi = dc + pad64(nc);
    float* a_lo = a_hi + pad64(a_split);
    float* b_hi = a_lo + pad64(a_split);
    float* b_lo = b_hi + pad64(b_split);
    int* flag = reinterpret_cast<int*>(b_lo + pad64(b_split));
    if (prepared) {
        db = const_cast<float*>(prepared->b);
        if (!fused) { b_hi = prepared->split; b_lo = prepared->split + n * kpad; }
    }

    // Everything that can fail runs inside `enqueue`; whatever it returns, the three streams are drained before the
    // borrowed host slices go back to the caller and the scratch is released (no leak, no copy still in flight).
    std::vector<Event> up(units + 1), done(units), down(pipe_trace() ? units : 0);
    Event ready, drained;
    auto enqueue = [&]() -> int {
    if (prepared && prepared->flag) TRN_CUDA(cudaMemcpyAsync(flag, prepared->flag, sizeof(int), cudaMemcpyDeviceToDevice, s_main));
    else TRN_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s_main));
    for (auto& e : down) TRN_TRY(e.create());

    TRN_TRY(ready.create());
    TRN_TRY(drained.create());
    for (auto& e : up) TRN_TRY(e.create());
    for (auto& e : done) TRN_TRY(e.create());
    // the allocation (stream-ordered on s_main) must be complete before the copy stream writes into it
    TRN_CUDA(cudaEventRecord(ready.e, s_main));
    TRN_CUDA(cudaStreamWaitEvent(s_up, ready.e, 0));
    TRN_CUDA(cudaStreamWaitEvent(s_down, ready.e, 0));

    int st = TRN_OK;
    if (single && !prepared) {
        TRN_CUDA(cudaMemcpyAsync(db, b, nb * sizeof(float), cudaMemcpyHostToDevice, s_up));
        TRN_CUDA(cudaEventRecord(up[units].e, s_up));
        TRN_CUDA(cudaStreamWaitEvent(s_main, up[units].e, 0));
        if (!fused) st = gemm_tc_split_b(db, b_hi, b_lo, 1, k, n, flag, s_main);
    }
    for (size_t u = 0; u < units && st == TRN_OK; ++u) {
        if (single) {
            const size_t r0 = row_start[u], rows = row_start[u + 1] - r0;
            cudaStream_t s_c = s_main;
            TRN_CUDA(cudaMemcpyAsync(da + r0 * k, a + r0 * k, rows * k * sizeof(float), cudaMemcpyHostToDevice, s_up));
            TRN_CUDA(cudaEventRecord(up[u].e, s_up));
            TRN_CUDA(cudaStreamWaitEvent(s_c, up[u].e, 0));
            if (fused) {
                st = gemm_tc_fused_main(da + r0 * k, db, dc + r0 * n, 1, rows, k, n, flag, s_c);
            } else {
                st = gemm_tc_split_a(da + r0 * k, a_hi + r0 * kpad, a_lo + r0 * kpad, 1, rows, k, flag, s_c);
                if (st == TRN_OK) st = gemm_tc_main(a_hi + r0 * kpad, a_lo + r0 * kpad, b_hi, b_lo, dc + r0 * n, 1, rows, k, n, 3, flag, s_c);
            }
            if (st == TRN_OK) st = launch_gemm_simt(da + r0 * k, db, dc + r0 * n, 1, rows, k, n, s_c, flag);
            TRN_CUDA(cudaEventRecord(done[u].e, s_c));
            TRN_CUDA(cudaStreamWaitEvent(s_down, done[u].e, 0));
            TRN_CUDA(cudaMemcpyAsync(c + r0 * n, dc + r0 * n, rows * n * sizeof(float), cudaMemcpyDeviceToHost, s_down));
            if (pipe_trace()) TRN_CUDA(cudaEventRecord(down[u].e, s_down));
        } else {
            const size_t b0 = u * per_block, cnt = batch - b0 < per_block ? batch - b0 : per_block;
            TRN_CUDA(cudaMemcpyAsync(da + b0 * m * k, a + b0 * m * k, cnt * m * k * sizeof(float), cudaMemcpyHostToDevice, s_up));
            TRN_CUDA(cudaMemcpyAsync(db + b0 * k * n, b + b0 * k * n, cnt * k * n * sizeof(float), cudaMemcpyHostToDevice, s_up));
            TRN_CUDA(cudaEventRecord(up[u].e, s_up));
            TRN_CUDA(cudaStreamWaitEvent(s_main, up[u].e, 0));
            if (fused) {
                st = gemm_tc_fused_main(da + b0 * m * k, db + b0 * k * n, dc + b0 * m * n, cnt, m, k, n, flag, s_main);
            } else {
                st = gemm_tc_split_a(da + b0 * m * k, a_hi + b0 * m * kpad, a_lo + b0 * m * kpad, cnt, m, k, flag, s_main);
                if (st == TRN_OK) st = gemm_tc_split_b(db + b0 * k * n, b_hi + b0 * n * kpad, b_lo + b0 * n * kpad, cnt, k, n, flag, s_main);
                if (st == TRN_OK) st = gemm_tc_main(a_hi + b0 * m * kpad, a_lo + b0 * m * kpad, b_hi + b0 * n * kpad, b_lo + b0 * n * kpad,
                                                    dc + b0 * m * n, cnt, m, k, n, 3, flag, s_main);
            }
            if (st == TRN_OK) st = launch_gemm_simt(da + b0 * m * k, db + b0 * k * n, dc + b0 * m * n, cnt, m, k, n, s_main, flag);
            TRN_CUDA(cudaEventRecord(done[u].e, s_main));
            TRN_CUDA(cudaStreamWaitEvent(s_down, done[u].e, 0));
            TRN_CUDA(cudaMemcpyAsync(c + b0 * m * n, dc + b0 * m * n, cnt * m * n * sizeof(float), cudaMemcpyDeviceToHost, s_down));
        }
    }
    return st;
    };
    const int st = enqueue();
    // the host slices are borrowed: everything must have landed before the call returns
    cudaError_t e1 = cudaStreamSynchronize(s_up), e2 = cudaStreamSynchronize(s_main), e3 = cudaStreamSynchronize(s_down);
    if (pipe_trace() && single && st == TRN_OK && e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess) {
        auto ms = [&](const Event& x) { float t = 0.f; cudaEventElapsedTime(&t, ready.e, x.e); return t; };
        fprintf(stderr, "[pipe trace] %zu blocks:", units);
        if (!prepared) fprintf(stderr, " B up %.2f |", ms(up[units]));
        for (size_t u = 0; u < units; ++u) fprintf(stderr, " [%zu rows] up %.2f gemm %.2f down %.2f |", row_start[u + 1] - row_start[u], ms(up[u]), ms(done[u]), ms(down[u]));
        fprintf(stderr, "\n");
    }
    scratch_free(dev, s_main);
    if (st != TRN_OK) return st;
    TRN_CUDA(e1);
    TRN_CUDA(e2);
    TRN_CUDA(e3);
    return TRN_OK;
}

static int host_gemm(const float* a, const float* b, float* c, size_t batch, size_t m, size_t k, size_t n) {
    TRN_HOST_LOCK();
    Context* cx = ctx();
    const size_t na = batch * m * k, nb = batch * k * n, nc = batch * m * n;
    if (nc == 0) return TRN_OK;
    // large tensor-core products from pinned host memory: overlap the PCIe transfers with the math
    if (g_engine.load() == 0 && k > 0 && m > 1 && gemm_auto_uses_tc(m, k, n) && (na + nb + nc) * sizeof(float) >= ((size_t)64 << 20) &&
        (batch > 1 || m >= 1024) && is_pinned_host(a) && is_pinned_host(b) && is_pinned_host(c))
        return host_gemm_pipelined(a, b, c, batch, m, k, n, nullptr);
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
    TRN_HOST_LOCK();
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
    TRN_HOST_LOCK();
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

// ===================== remaining VectorBackend surface =====================
// Validation per op follows src/vector.rs: sub/div/lerp/fma -> SizeMismatch; relu/swish/tanh on an empty vector ->
// EmptyVector (:1670, :2293, :3950); the other maps return an empty result for an empty input (:2827-4182);
// clamp with min > max -> InvalidInput("Invalid clamp range: min ({}) > max ({})") (:2964-2969).
#define TRN_UNARY(NAME, OP, EMPTY_IS_ERROR)                                                              \
    int trn_##NAME##_f32_dev(const float* a, size_t n, float* out, void* stream) {                       \
        if (EMPTY_IS_ERROR) TRN_TRY(check_nonempty_emptyvec(n));                                         \
        TRN_TRY(need_ctx());                                                                             \
        return launch_map(Map::OP, a, nullptr, nullptr, out, n, 0.f, 0.f, resolve_stream(stream));       \
    }                                                                                                    \
    int trn_##NAME##_f32(const float* a, size_t n, float* out) {                                         \
        if (EMPTY_IS_ERROR) TRN_TRY(check_nonempty_emptyvec(n));                                         \
        TRN_TRY(need_ctx());                                                                             \
        return host_map(Map::OP, a, nullptr, nullptr, out, n);                                           \
    }
TRN_UNARY(abs, Abs, false)
TRN_UNARY(relu, Relu, true)
TRN_UNARY(exp, Exp, false)
TRN_UNARY(swish, Swish, true)
TRN_UNARY(tanh, Tanh, true)
TRN_UNARY(sqrt, Sqrt, false)
TRN_UNARY(recip, Recip, false)
TRN_UNARY(ln, Ln, false)
TRN_UNARY(log2, Log2, false)
TRN_UNARY(log10, Log10, false)
TRN_UNARY(sin, Sin, false)
TRN_UNARY(cos, Cos, false)
TRN_UNARY(tan, Tan, false)
TRN_UNARY(floor, Floor, false)
TRN_UNARY(ceil, Ceil, false)
TRN_UNARY(round, Round, false)
#undef TRN_UNARY

#define TRN_BINARY(NAME, OP)                                                                             \
    int trn_##NAME##_f32_dev(const float* a, size_t na, const float* b, size_t nb, float* out, void* stream) { \
        TRN_TRY(check_same_len(na, nb));                                                                 \
        TRN_TRY(need_ctx());                                                                             \
        return launch_map(Map::OP, a, b, nullptr, out, na, 0.f, 0.f, resolve_stream(stream));            \
    }                                                                                                    \
    int trn_##NAME##_f32(const float* a, size_t na, const float* b, size_t nb, float* out) {             \
        TRN_TRY(check_same_len(na, nb));                                                                 \
        TRN_TRY(need_ctx());                                                                             \
        return host_map(Map::OP, a, b, nullptr, out, na);                                                \
    }
TRN_BINARY(sub, Sub)
TRN_BINARY(div, Div)
#undef TRN_BINARY

int trn_scale_f32_dev(const float* a, size_t n, float scalar, float* out, void* stream) {
    TRN_TRY(need_ctx());
    return launch_map(Map::Scale, a, nullptr, nullptr, out, n, scalar, 0.f, resolve_stream(stream));
}
int trn_scale_f32(const float* a, size_t n, float scalar, float* out) {
    TRN_TRY(need_ctx());
    return host_map(Map::Scale, a, nullptr, nullptr, out, n, scalar);
}
int trn_clamp_f32_dev(const float* a, size_t n, float min_val, float max_val, float* out, void* stream) {
    TRN_TRY(check_clamp(min_val, max_val));
    TRN_TRY(need_ctx());
    return launch_map(Map::Clamp, a, nullptr, nullptr, out, n, min_val, max_val, resolve_stream(stream));
}
int trn_clamp_f32(const float* a, size_t n, float min_val, float max_val, float* out) {
    TRN_TRY(check_clamp(min_val, max_val));
    TRN_TRY(need_ctx());
    return host_map(Map::Clamp, a, nullptr, nullptr, out, n, min_val, max_val);
}
int trn_lerp_f32_dev(const float* a, size_t na, const float* b, size_t nb, float t, float* out, void* stream) {
    TRN_TRY(check_same_len(na, nb));
    TRN_TRY(need_ctx());
    return launch_map(Map::Lerp, a, b, nullptr, out, na, t, 0.f, resolve_stream(stream));
}
int trn_lerp_f32(const float* a, size_t na, const float* b, size_t nb, float t, float* out) {
    TRN_TRY(check_same_len(na, nb));
    TRN_TRY(need_ctx());
    return host_map(Map::Lerp, a, b, nullptr, out, na, t);
}
int trn_fma_f32_dev(const float* a, size_t na, const float* b, size_t nb, const float* c, size_t nc, float* out, void* stream) {
    TRN_TRY(check_same_len(na, nb));
    TRN_TRY(check_same_len(na, nc));
    TRN_TRY(need_ctx());
    return launch_map(Map::Fma, a, b, c, out, na, 0.f, 0.f, resolve_stream(stream));
}
int trn_fma_f32(const float* a, size_t na, const float* b, size_t nb, const float* c, size_t nc, float* out) {
    TRN_TRY(check_same_len(na, nb));
    TRN_TRY(check_same_len(na, nc));
    TRN_TRY(need_ctx());
    return host_map(Map::Fma, a, b, c, out, na);
}

// ===================== the rest of Vector's element-wise / statistics API (src/vector.rs) =====================
// Validation per op follows src/vector.rs: hardswish / mish / selu / leaky_relu / elu on an empty vector -> EmptyVector
// (:2410, :2478, :2547, :1981, :2086); the plain maps (neg ... atanh, pow, clip) return an empty result for an empty
// input; minimum / maximum / copysign -> SizeMismatch (:4329, :4365, :4293).
#define TRN_UNARY(NAME, OP, EMPTY_IS_ERROR)                                                              \
    int trn_##NAME##_f32_dev(const float* a, size_t n, float* out, void* stream) {                       \
        if (EMPTY_IS_ERROR) TRN_TRY(check_nonempty_emptyvec(n));                                         \
        TRN_TRY(need_ctx());                                                                             \
        return launch_map(Map::OP, a, nullptr, nullptr, out, n, 0.f, 0.f, resolve_stream(stream));       \
    }                                                                                                    \
    int trn_##NAME##_f32(const float* a, size_t n, float* out) {                                         \
        if (EMPTY_IS_ERROR) TRN_TRY(check_nonempty_emptyvec(n));                                         \
        TRN_TRY(need_ctx());                                                                             \
        return host_map(Map::OP, a, nullptr, nullptr, out, n);                                           \
    }
TRN_UNARY(neg, Neg, false)
TRN_UNARY(signum, Signum, false)
TRN_UNARY(trunc, Trunc, false)
TRN_UNARY(fract, Fract, false)
TRN_UNARY(sinh, Sinh, false)
TRN_UNARY(cosh, Cosh, false)
TRN_UNARY(asin, Asin, false)
TRN_UNARY(acos, Acos, false)
TRN_UNARY(atan, Atan, false)
TRN_UNARY(asinh, Asinh, false)
TRN_UNARY(acosh, Acosh, false)
TRN_UNARY(atanh, Atanh, false)
TRN_UNARY(hardswish, Hardswish, true)
TRN_UNARY(mish, Mish, true)
TRN_UNARY(selu, Selu, true)
#undef TRN_UNARY

#define TRN_BINARY(NAME, OP)                                                                             \
    int trn_##NAME##_f32_dev(const float* a, size_t na, const float* b, size_t nb, float* out, void* stream) { \
        TRN_TRY(check_same_len(na, nb));                                                                 \
        TRN_TRY(need_ctx());                                                                             \
        return launch_map(Map::OP, a, b, nullptr, out, na, 0.f, 0.f, resolve_stream(stream));            \
    }                                                                                                    \
    int trn_##NAME##_f32(const float* a, size_t na, const float* b, size_t nb, float* out) {             \
        TRN_TRY(check_same_len(na, nb));                                                                 \
        TRN_TRY(need_ctx());                                                                             \
        return host_map(Map::OP, a, b, nullptr, out, na);                                                \
    }
TRN_BINARY(minimum, Minimum)
TRN_BINARY(maximum, Maximum)
TRN_BINARY(copysign, Copysign)
#undef TRN_BINARY

static int check_leaky_relu(size_t n, float slope) {   // src/vector.rs:1981-1991
    TRN_TRY(check_nonempty_emptyvec(n));
    if (!(slope >= 0.0f && slope < 1.0f))
        return fail(TRN_INVALID_INPUT, "negative_slope must be in [0.0, 1.0), got %s", fmt_f32(slope).c_str());
    return TRN_OK;
}
static int check_elu(size_t n, float alpha) {          // src/vector.rs:2086-2096
    TRN_TRY(check_nonempty_emptyvec(n));
    if (alpha <= 0.0f) return fail(TRN_INVALID_INPUT, "alpha must be > 0, got %s", fmt_f32(alpha).c_str());
    return TRN_OK;
}
static int check_clip(float lo, float hi) {            // src/vector.rs:1449-1454
    if (lo > hi)
        return fail(TRN_INVALID_INPUT, "min_val (%s) must be <= max_val (%s)", fmt_f32(lo).c_str(), fmt_f32(hi).c_str());
    return TRN_OK;
}
int trn_leaky_relu_f32_dev(const float* a, size_t n, float negative_slope, float* out, void* stream) {
    TRN_TRY(check_leaky_relu(n, negative_slope));
    TRN_TRY(need_ctx());
    return launch_map(Map::LeakyRelu, a, nullptr, nullptr, out, n, negative_slope, 0.f, resolve_stream(stream));
}
int trn_leaky_relu_f32(const float* a, size_t n, float negative_slope, float* out) {
    TRN_TRY(check_leaky_relu(n, negative_slope));
    TRN_TRY(need_ctx());
    return host_map(Map::LeakyRelu, a, nullptr, nullptr, out, n, negative_slope);
}
int trn_elu_f32_dev(const float* a, size_t n, float alpha, float* out, void* stream) {
    TRN_TRY(check_elu(n, alpha));
    TRN_TRY(need_ctx());
    return launch_map(Map::Elu, a, nullptr, nullptr, out, n, alpha, 0.f, resolve_stream(stream));
}
int trn_elu_f32(const float* a, size_t n, float alpha, float* out) {
    TRN_TRY(check_elu(n, alpha));
    TRN_TRY(need_ctx());
    return host_map(Map::Elu, a, nullptr, nullptr, out, n, alpha);
}
int trn_pow_f32_dev(const float* a, size_t n, float exponent, float* out, void* stream) {
    TRN_TRY(need_ctx());
    return launch_map(Map::Pow, a, nullptr, nullptr, out, n, exponent, 0.f, resolve_stream(stream));
}
int trn_pow_f32(const float* a, size_t n, float exponent, float* out) {
    TRN_TRY(need_ctx());
    return host_map(Map::Pow, a, nullptr, nullptr, out, n, exponent);
}
// Vector::clip (src/vector.rs:1448): x.max(min).min(max) — clamp's arithmetic, its own error text
int trn_clip_f32_dev(const float* a, size_t n, float min_val, float max_val, float* out, void* stream) {
    TRN_TRY(check_clip(min_val, max_val));
    TRN_TRY(need_ctx());
    return launch_map(Map::Clamp, a, nullptr, nullptr, out, n, min_val, max_val, resolve_stream(stream));
}
int trn_clip_f32(const float* a, size_t n, float min_val, float max_val, float* out) {
    TRN_TRY(check_clip(min_val, max_val));
    TRN_TRY(need_ctx());
    return host_map(Map::Clamp, a, nullptr, nullptr, out, n, min_val, max_val);
}
// out = (a - shift) * scale, the second half of zscore / minmax_normalize for device-resident callers
int trn_affine_f32_dev(const float* a, size_t n, float shift, float scale, float* out, void* stream) {
    TRN_TRY(need_ctx());
    return launch_map(Map::Affine, a, nullptr, nullptr, out, n, shift, scale, resolve_stream(stream));
}

// ---- reductions: sum_kahan, norm_l1, norm_linf (empty -> 0, src/vector.rs:848, :2708, :2770) ----------------
#define TRN_REDUCE(NAME, OP)                                                                             \
    int trn_##NAME##_f32_dev(const float* a, size_t n, float* out, void* stream) {                       \
        TRN_TRY(need_ctx());                                                                             \
        return launch_reduce(Reduce::OP, a, nullptr, n, out, resolve_stream(stream));                    \
    }                                                                                                    \
    int trn_##NAME##_f32(const float* a, size_t n, float* out) {                                         \
        TRN_TRY(need_ctx());                                                                             \
        return host_reduce_f32(a, n, nullptr, 0, out, [&](const float* da, const float*, float* o, cudaStream_t s) { \
            return launch_reduce(Reduce::OP, da, nullptr, n, o, s);                                      \
        });                                                                                              \
    }
TRN_REDUCE(sum_kahan, SumKahan)
TRN_REDUCE(norm_l1, SumAbs)
TRN_REDUCE(norm_linf, MaxAbs)
#undef TRN_REDUCE

// mean / variance / stddev (src/vector.rs:935-1030): empty -> EmptyVector; mean = sum / n;
// variance = E[x^2] - mean^2 with both moments from the device reductions; stddev = sqrt(variance)
static int host_moments(const float* a, size_t n, float* mean, float* var) {
    TRN_HOST_LOCK();
    TRN_TRY(check_nonempty_emptyvec(n));
    TRN_TRY(need_ctx());
    Context* c = ctx();
    Workspace* w = workspace(c->stream);
    if (!w) return fail(TRN_GPU_ERROR, "failed to allocate the reduction workspace");
    DevTemp da(c->stream);
    TRN_TRY(da.alloc(n));
    TRN_TRY(upload(da.p, a, n, c->stream));
    float sum = 0.f, sumsq = 0.f;
    TRN_TRY(launch_reduce(Reduce::Sum, da.p, nullptr, n, w->scalar_f32, c->stream));
    TRN_CUDA(cudaMemcpyAsync(w->host_f32, w->scalar_f32, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    TRN_CUDA(cudaStreamSynchronize(c->stream));
    sum = *w->host_f32;
    if (var) {
        TRN_TRY(launch_reduce(Reduce::SumSq, da.p, nullptr, n, w->scalar_f32, c->stream));
        TRN_CUDA(cudaMemcpyAsync(w->host_f32, w->scalar_f32, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        TRN_CUDA(cudaStreamSynchronize(c->stream));
        sumsq = *w->host_f32;
    }
    const float m = sum / (float)n;
    if (mean) *mean = m;
    if (var) *var = sumsq / (float)n - m * m;
    return TRN_OK;
}
int trn_mean_f32(const float* a, size_t n, float* out) { return host_moments(a, n, out, nullptr); }
int trn_variance_f32(const float* a, size_t n, float* out) { return host_moments(a, n, nullptr, out); }
int trn_stddev_f32(const float* a, size_t n, float* out) {
    float v = 0.f;
    TRN_TRY(host_moments(a, n, nullptr, &v));
    *out = sqrtf(v);
    return TRN_OK;
}

// ---- statistics composed in Vector itself (src/vector.rs:898-1290, :1386): ONE upload, the reductions and the
//      map on the resident copy, one download.  Scalars come back to the host between the steps, as in the reference.
namespace {
struct Resident {          // a host vector uploaded once, with scalar reductions read back to the host
    Context* c;
    Workspace* w;
    DevTemp d;
    size_t n;
    explicit Resident(Context* c_) : c(c_), w(workspace(c_->stream)), d(c_->stream), n(0) {}
    int put(const float* a, size_t n_) {
        if (!w) return fail(TRN_GPU_ERROR, "failed to allocate the reduction workspace");
        n = n_;
        TRN_TRY(d.alloc(n));
        return upload(d.p, a, n, c->stream);
    }
    int scalar(float* out) {
        TRN_CUDA(cudaMemcpyAsync(w->host_f32, w->scalar_f32, sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        TRN_CUDA(cudaStreamSynchronize(c->stream));
        *out = *w->host_f32;
        return TRN_OK;
    }
    int reduce(Reduce op, const float* other, float* out) {
        TRN_TRY(launch_reduce(op, d.p, other, n, w->scalar_f32, c->stream));
        return scalar(out);
    }
    int extremum(int is_max, float* out) {
        TRN_TRY(launch_argreduce(is_max, d.p, n, nullptr, w->scalar_f32, c->stream));
        return scalar(out);
    }
    // mean = sum / n; variance = E[x^2] - mean^2 (src/vector.rs:935-1015)
    int moments(float* mean, float* var) {
        float sum = 0.f, sumsq = 0.f;
        TRN_TRY(reduce(Reduce::Sum, nullptr, &sum));
        const float m = sum / (float)n;
        if (mean) *mean = m;
        if (var) {
            TRN_TRY(reduce(Reduce::SumSq, nullptr, &sumsq));
            *var = sumsq / (float)n - m * m;
        }
        return TRN_OK;
    }
    int affine_out(float shift, float scale, float* out) {
        DevTemp o(c->stream);
        TRN_TRY(o.alloc(n));
        TRN_TRY(launch_map(Map::Affine, d.p, nullptr, nullptr, o.p, n, shift, scale, c->stream));
        return download(out, o.p, n, c->stream);
    }
};
}  // namespace

// Vector::sum_of_squares (src/vector.rs:898): dot(self, self); empty -> 0
int trn_sum_of_squares_f32(const float* a, size_t n, float* out) {
    TRN_TRY(need_ctx());
    if (n == 0) { *out = 0.f; return TRN_OK; }
    TRN_HOST_LOCK();
    Resident x(ctx());
    TRN_TRY(x.put(a, n));
    return x.reduce(Reduce::SumSq, nullptr, out);
}
// Vector::covariance (src/vector.rs:1063): E[xy] - mean_x * mean_y; empty -> EmptyVector, then SizeMismatch
static int covariance_resident(Resident& x, Resident& y, float* out) {
    float mx = 0.f, my = 0.f, dot = 0.f;
    TRN_TRY(x.moments(&mx, nullptr));
    TRN_TRY(y.moments(&my, nullptr));
    TRN_TRY(x.reduce(Reduce::Dot, y.d.p, &dot));
    *out = dot / (float)x.n - mx * my;
    return TRN_OK;
}
int trn_covariance_f32(const float* a, size_t na, const float* b, size_t nb, float* out) {
    TRN_TRY(check_nonempty_emptyvec(na));
    TRN_TRY(check_same_len(na, nb));
    TRN_TRY(need_ctx());
    TRN_HOST_LOCK();
    Resident x(ctx()), y(ctx());
    TRN_TRY(x.put(a, na));
    TRN_TRY(y.put(b, nb));
    return covariance_resident(x, y, out);
}
// Vector::correlation (src/vector.rs:1119): cov / (std_x * std_y) clamped to [-1, 1]; |std| < 1e-10 -> DivisionByZero
int trn_correlation_f32(const float* a, size_t na, const float* b, size_t nb, float* out) {
    TRN_TRY(check_nonempty_emptyvec(na));
    TRN_TRY(check_same_len(na, nb));
    TRN_TRY(need_ctx());
    TRN_HOST_LOCK();
    Resident x(ctx()), y(ctx());
    TRN_TRY(x.put(a, na));
    TRN_TRY(y.put(b, nb));
    float cov = 0.f, vx = 0.f, vy = 0.f;
    TRN_TRY(covariance_resident(x, y, &cov));
    TRN_TRY(x.moments(nullptr, &vx));
    TRN_TRY(y.moments(nullptr, &vy));
    const float sx = sqrtf(vx), sy = sqrtf(vy);
    if (fabsf(sx) < 1e-10f || fabsf(sy) < 1e-10f) return fail(TRN_DIVISION_BY_ZERO, "Division by zero");
    const float corr = cov / (sx * sy);
    *out = corr != corr ? corr : fminf(fmaxf(corr, -1.0f), 1.0f);   // f32::clamp (src/vector.rs:1153) propagates NaN
    return TRN_OK;
}
// Vector::zscore (src/vector.rs:1180): (x - mean) * (1 / stddev); empty -> EmptyVector; |stddev| < 1e-10 -> DivisionByZero
int trn_zscore_f32(const float* a, size_t n, float* out) {
    TRN_TRY(check_nonempty_emptyvec(n));
    TRN_TRY(need_ctx());
    TRN_HOST_LOCK();
    Resident x(ctx());
    TRN_TRY(x.put(a, n));
    float mean = 0.f, var = 0.f;
    TRN_TRY(x.moments(&mean, &var));
    const float sd = sqrtf(var);
    if (fabsf(sd) < 1e-10f) return fail(TRN_DIVISION_BY_ZERO, "Division by zero");
    return x.affine_out(mean, 1.0f / sd, out);
}
// Vector::minmax_normalize (src/vector.rs:1248): (x - min) * (1 / (max - min)); |range| < 1e-10 -> DivisionByZero.
// min / max are exact and the map is two rounded operations -> bit-exact against the reference.
int trn_minmax_normalize_f32(const float* a, size_t n, float* out) {
    TRN_TRY(check_nonempty_emptyvec(n));
    TRN_TRY(need_ctx());
    TRN_HOST_LOCK();
    Resident x(ctx());
    TRN_TRY(x.put(a, n));
    float lo = 0.f, hi = 0.f;
    TRN_TRY(x.extremum(0, &lo));
    TRN_TRY(x.extremum(1, &hi));
    const float range = hi - lo;
    if (fabsf(range) < 1e-10f) return fail(TRN_DIVISION_BY_ZERO, "Division by zero");
    return x.affine_out(lo, 1.0f / range, out);
}
// Vector::layer_norm_simple (src/vector.rs:1386): layer_norm without gamma / beta: (x - mean) * inv_std with
// variance = sum((x - mean)^2) / n — the row kernels with gamma == beta == nullptr
int trn_layer_norm_simple_rows_f32_dev(const float* a, float eps, float* out, size_t rows, size_t cols, void* stream) {
    TRN_TRY(check_nonempty_emptyvec(rows * cols));
    TRN_TRY(need_ctx());
    return launch_layer_norm_rows(a, nullptr, nullptr, eps, out, rows, cols, resolve_stream(stream));
}
int trn_layer_norm_simple_rows_f32(const float* a, float eps, float* out, size_t rows, size_t cols) {
    TRN_TRY(check_nonempty_emptyvec(rows * cols));
    TRN_TRY(need_ctx());
    TRN_HOST_LOCK();
    Context* c = ctx();
    const size_t n = rows * cols;
    DevTemp da(c->stream), dout(c->stream);
    TRN_TRY(da.alloc(n));
    TRN_TRY(dout.alloc(n));
    TRN_TRY(upload(da.p, a, n, c->stream));
    TRN_TRY(launch_layer_norm_rows(da.p, nullptr, nullptr, eps, dout.p, rows, cols, c->stream));
    return download(out, dout.p, n, c->stream);
}

// ===================== callers next to the path =====================
// Matrix::vecmat (src/matrix.rs:1782): y = v^T A; v_len != rows -> InvalidInput
int trn_vecmat_f32_dev(const float* v, size_t v_len, const float* a, size_t rows, size_t cols, float* y, void* stream) {
    if (v_len != rows)
        return fail(TRN_INVALID_INPUT, "Vector length %zu does not match matrix rows %zu for vector-matrix multiplication",
                    v_len, rows);
    TRN_TRY(need_ctx());
    cudaStream_t s = resolve_stream(stream);
    if (rows == 0 && cols > 0) {
        TRN_CUDA(cudaMemsetAsync(y, 0, cols * sizeof(float), s));
        return TRN_OK;
    }
    return launch_vecmat(v, a, rows, cols, y, s, false);
}
int trn_vecmat_f32(const float* v, size_t v_len, const float* a, size_t rows, size_t cols, float* y) {
    if (v_len != rows)
        return fail(TRN_INVALID_INPUT, "Vector length %zu does not match matrix rows %zu for vector-matrix multiplication",
                    v_len, rows);
    TRN_TRY(need_ctx());
    TRN_HOST_LOCK();
    Context* c = ctx();
    if (cols == 0) return TRN_OK;
    DevTemp da(c->stream), dv(c->stream), dy(c->stream);
    TRN_TRY(da.alloc(rows * cols));
    TRN_TRY(dv.alloc(rows));
    TRN_TRY(dy.alloc(cols));
    TRN_TRY(upload(da.p, a, rows * cols, c->stream));
    TRN_TRY(upload(dv.p, v, rows, c->stream));
    if (rows == 0) TRN_CUDA(cudaMemsetAsync(dy.p, 0, cols * sizeof(float), c->stream));
    else TRN_TRY(launch_vecmat(dv.p, da.p, rows, cols, dy.p, c->stream, false));
    return download(y, dy.p, cols, c->stream);
}

// Matrix::embedding_lookup (src/matrix.rs:2008): out[r, :] = table[indices[r], :].  An index >= rows ->
// InvalidInput("Index {} at position {} is out of bounds for embedding table with {} rows") from the host-slice call
// (indices are host memory there); the resident twin cannot see its indices and writes a zero row instead.
int trn_embedding_lookup_f32_dev(const float* table, size_t rows, size_t cols, const uint64_t* indices, size_t n_indices,
                                 float* out, void* stream) {
    TRN_TRY(need_ctx());
    return launch_gather_rows(table, rows, cols, indices, n_indices, out, resolve_stream(stream));
}
int trn_embedding_lookup_f32(const float* table, size_t rows, size_t cols, const uint64_t* indices, size_t n_indices,
                             float* out) {
    for (size_t i = 0; i < n_indices; ++i)
        if (indices[i] >= rows)
            return fail(TRN_INVALID_INPUT, "Index %llu at position %zu is out of bounds for embedding table with %zu rows",
                        (unsigned long long)indices[i], i, rows);
    TRN_TRY(need_ctx());
    if (n_indices == 0 || cols == 0) return TRN_OK;
    TRN_HOST_LOCK();
    Context* c = ctx();
    DevTemp dt(c->stream), dout(c->stream), didx(c->stream);
    TRN_TRY(dt.alloc(rows * cols));
    TRN_TRY(dout.alloc(n_indices * cols));
    TRN_TRY(didx.alloc(2 * n_indices));   // u64 indices in a float-typed scratch block
    TRN_TRY(upload(dt.p, table, rows * cols, c->stream));
    TRN_CUDA(cudaMemcpyAsync(didx.p, indices, n_indices * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    TRN_TRY(launch_gather_rows(dt.p, rows, cols, reinterpret_cast<const uint64_t*>(didx.p), n_indices, dout.p, c->stream));
    return download(out, dout.p, n_indices * cols, c->stream);
}

// Vector::layer_norm (src/vector.rs:1316): `rows` vectors of `cols` elements sharing gamma / beta (rows == 1 is
// exactly the reference call).  Empty -> EmptyVector; gamma / beta length != cols -> SizeMismatch{cols, len}.
static int check_layer_norm(size_t rows, size_t cols, size_t ng, size_t nb) {
    TRN_TRY(check_nonempty_emptyvec(rows * cols));
    if (ng != cols) return fail_mismatch(cols, ng);
    if (nb != cols) return fail_mismatch(cols, nb);
    return TRN_OK;
}
int trn_layer_norm_rows_f32_dev(const float* a, const float* gamma, size_t gamma_len, const float* beta, size_t beta_len,
                                float eps, float* out, size_t rows, size_t cols, void* stream) {
    TRN_TRY(check_layer_norm(rows, cols, gamma_len, beta_len));
    TRN_TRY(need_ctx());
    return launch_layer_norm_rows(a, gamma, beta, eps, out, rows, cols, resolve_stream(stream));
}
int trn_layer_norm_rows_f32(const float* a, const float* gamma, size_t gamma_len, const float* beta, size_t beta_len,
                            float eps, float* out, size_t rows, size_t cols) {
    TRN_TRY(check_layer_norm(rows, cols, gamma_len, beta_len));
    TRN_TRY(need_ctx());
    TRN_HOST_LOCK();
    Context* c = ctx();
    const size_t n = rows * cols;
    DevTemp da(c->stream), dg(c->stream), db(c->stream), dout(c->stream);
    TRN_TRY(da.alloc(n));
    TRN_TRY(dg.alloc(cols));
    TRN_TRY(db.alloc(cols));
    TRN_TRY(dout.alloc(n));
    TRN_TRY(upload(da.p, a, n, c->stream));
    TRN_TRY(upload(dg.p, gamma, cols, c->stream));
    TRN_TRY(upload(db.p, beta, cols, c->stream));
    TRN_TRY(launch_layer_norm_rows(da.p, dg.p, db.p, eps, dout.p, rows, cols, c->stream));
    return download(out, dout.p, n, c->stream);
}

// Matrix::convolve2d (src/matrix.rs:1868): valid padding; kernel larger than the input -> InvalidInput
static int check_conv(size_t rows, size_t cols, size_t kr, size_t kc) {   // src/matrix.rs:1870-1875
    if (kr > rows || kc > cols)
        return fail(TRN_INVALID_INPUT, "Kernel size (%zux%zu) larger than input (%zux%zu)", kr, kc, rows, cols);
    return TRN_OK;
}
int trn_convolve2d_f32_dev(const float* in, size_t rows, size_t cols, const float* kernel, size_t k_rows, size_t k_cols,
                           float* out, void* stream) {
    TRN_TRY(check_conv(rows, cols, k_rows, k_cols));
    TRN_TRY(need_ctx());
    return launch_convolve2d(in, rows, cols, kernel, k_rows, k_cols, out, resolve_stream(stream));
}
int trn_convolve2d_f32(const float* in, size_t rows, size_t cols, const float* kernel, size_t k_rows, size_t k_cols, float* out) {
    TRN_TRY(check_conv(rows, cols, k_rows, k_cols));
    TRN_TRY(need_ctx());
    TRN_HOST_LOCK();
    Context* c = ctx();
    const size_t n_out = (rows - k_rows + 1) * (cols - k_cols + 1);
    if (n_out == 0 || k_rows * k_cols == 0) return TRN_OK;
    DevTemp din(c->stream), dk(c->stream), dout(c->stream);
    TRN_TRY(din.alloc(rows * cols));
    TRN_TRY(dk.alloc(k_rows * k_cols));
    TRN_TRY(dout.alloc(n_out));
    TRN_TRY(upload(din.p, in, rows * cols, c->stream));
    TRN_TRY(upload(dk.p, kernel, k_rows * k_cols, c->stream));
    TRN_TRY(launch_convolve2d(din.p, rows, cols, dk.p, k_rows, k_cols, dout.p, c->stream));
    return download(out, dout.p, n_out, c->stream);
}

// ---- fused attention (trueno-gpu/src/kernels/attention.rs:27-125) ---------------------------------------------
static int check_attention(size_t q_len, size_t k_len, size_t v_len, size_t heads, size_t seq, size_t d) {
    const size_t want = heads * seq * d;   // size checks in the style of src/matrix.rs:481-502
    const char* names[3] = {"Q", "K", "V"};
    const size_t lens[3] = {q_len, k_len, v_len};
    for (int i = 0; i < 3; ++i)
        if (lens[i] != want)
            return fail(TRN_INVALID_INPUT, "%s data size mismatch: expected %zu (%zu\xC3\x97%zu\xC3\x97%zu), got %zu", names[i], want,
                        heads, seq, d, lens[i]);
    if (d > attention_max_head_dim())
        return fail(TRN_INVALID_INPUT, "head_dim %zu exceeds the supported maximum %zu", d, attention_max_head_dim());
    return TRN_OK;
}
static int attention_engine() {
    const int e = g_engine.load();
    return e == 1 ? 1 : (e == 2 ? 2 : 0);
}
int trn_attention_f32_dev(const float* q, size_t q_len, const float* k, size_t k_len, const float* v, size_t v_len, float* out,
                          size_t heads, size_t seq_len, size_t head_dim, float scale, int causal, void* stream) {
    TRN_TRY(check_attention(q_len, k_len, v_len, heads, seq_len, head_dim));
    TRN_TRY(need_ctx());
    // engine 2 on a head_dim the tensor kernel cannot take falls back to auto rather than failing a forced test sweep
    const int engine = (attention_engine() == 2 && head_dim > 128) ? 0 : attention_engine();
    return launch_attention(q, k, v, out, heads, seq_len, head_dim, scale, causal, engine, resolve_stream(stream));
}
int trn_attention_f32(const float* q, size_t q_len, const float* k, size_t k_len, const float* v, size_t v_len, float* out,
                      size_t heads, size_t seq_len, size_t head_dim, float scale, int causal) {
    TRN_TRY(check_attention(q_len, k_len, v_len, heads, seq_len, head_dim));
    TRN_TRY(need_ctx());
    TRN_HOST_LOCK();
    Context* c = ctx();
    const size_t n = heads * seq_len * head_dim;
    if (n == 0) return TRN_OK;
    DevTemp dq(c->stream), dk(c->stream), dv(c->stream), dout(c->stream);
    TRN_TRY(dq.alloc(n));
    TRN_TRY(dk.alloc(n));
    TRN_TRY(dv.alloc(n));
    TRN_TRY(dout.alloc(n));
    TRN_TRY(upload(dq.p, q, n, c->stream));
    TRN_TRY(upload(dk.p, k, n, c->stream));
    TRN_TRY(upload(dv.p, v, n, c->stream));
    const int engine = (attention_engine() == 2 && head_dim > 128) ? 0 : attention_engine();
    TRN_TRY(launch_attention(dq.p, dk.p, dv.p, dout.p, heads, seq_len, head_dim, scale, causal, engine, c->stream));
    return download(out, dout.p, n, c->stream);
}

// ---- SymmetricEigen (src/eigen.rs:108-141) ----------------------------------------------------------------------
static int check_eigen(size_t rows, size_t cols) {
    if (rows != cols) return fail(TRN_INVALID_INPUT, "Matrix must be square for eigendecomposition, got %zux%zu", rows, cols);
    if (rows == 0) return fail(TRN_INVALID_INPUT, "Cannot compute eigendecomposition of empty matrix");
    if (rows > eigen_max_n())
        return fail(TRN_INVALID_INPUT, "matrix dimension %zu exceeds the supported maximum %zu", rows, eigen_max_n());
    return TRN_OK;
}
int trn_symmetric_eigen_f32_dev(const float* a, size_t rows, size_t cols, float* eigenvalues, float* eigenvectors, void* stream) {
    TRN_TRY(check_eigen(rows, cols));
    TRN_TRY(need_ctx());
    return launch_symmetric_eigen(a, rows, eigenvalues, eigenvectors, nullptr, resolve_stream(stream));
}
int trn_symmetric_eigen_f32(const float* a, size_t rows, size_t cols, float* eigenvalues, float* eigenvectors) {
    TRN_TRY(check_eigen(rows, cols));
    TRN_TRY(need_ctx());
    TRN_HOST_LOCK();
    Context* c = ctx();
    DevTemp da(c->stream), dvals(c->stream), dvecs(c->stream);
    TRN_TRY(da.alloc(rows * rows));
    TRN_TRY(dvals.alloc(rows));
    TRN_TRY(dvecs.alloc(rows * rows));
    TRN_TRY(upload(da.p, a, rows * rows, c->stream));
    TRN_TRY(launch_symmetric_eigen(da.p, rows, dvals.p, dvecs.p, nullptr, c->stream));
    TRN_TRY(download(eigenvalues, dvals.p, rows, c->stream));
    return download(eigenvectors, dvecs.p, rows * rows, c->stream);
}

}  // extern "C"
