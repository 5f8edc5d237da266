// context.cu — process-global device context, thread-local error state, per-stream workspaces,
// stream-ordered scratch, device buffers and pinned host staging.
//
// Replaces, for the CUDA path, what the reference spreads over GpuBackend/GpuDevice construction
// (src/backends/gpu/mod.rs:63-120, device.rs:31-80 — a fresh device per matmul call,
// src/matrix.rs:1549) with ONE lazily created context per process.  There is no CPU fallback:
// when no sm_100 device is usable, ctx() fails and every entry point reports GpuError.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <mutex>
#include <unordered_map>

#include "common.cuh"

namespace trn {

// ---- thread-local error state -------------------------------------------------------------------
static thread_local char t_msg[1024];
static thread_local uint64_t t_expected = 0, t_actual = 0;

int fail(int status, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_msg, sizeof t_msg, fmt, ap);
    va_end(ap);
    return status;
}
int fail_mismatch(size_t expected, size_t actual) {
    t_expected = expected;
    t_actual = actual;
    // Display text of TruenoError::SizeMismatch (src/error.rs:17-18)
    snprintf(t_msg, sizeof t_msg, "Size mismatch: expected %zu, got %zu", expected, actual);
    return TRN_SIZE_MISMATCH;
}
int fail_cuda(cudaError_t e, const char* what) {
    snprintf(t_msg, sizeof t_msg, "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    cudaGetLastError();  // clear the sticky flag for non-fatal errors
    return TRN_GPU_ERROR;
}

// ---- context --------------------------------------------------------------------------------------
static std::mutex g_mu;
static Context* g_ctx = nullptr;
static std::unordered_map<cudaStream_t, Workspace*> g_ws;
static std::atomic<uint64_t> g_launches{0};

// pinned staging ring for pageable host memory
constexpr size_t kStageBytes = 32u << 20;
static std::mutex g_stage_mu;
static char* g_stage[2] = {nullptr, nullptr};
static cudaEvent_t g_stage_ev[2];

void count_launch(unsigned n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
void uncount_launch(unsigned n) { g_launches.fetch_sub(n, std::memory_order_relaxed); }

static int init_locked(int device) {
    if (g_ctx) {
        if (device >= 0 && device != g_ctx->device)
            return fail(TRN_GPU_ERROR, "trn_cuda_init(%d): process is already bound to device %d", device,
                        g_ctx->device);
        return TRN_OK;
    }
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(TRN_GPU_ERROR, "no CUDA device available (%s); the CUDA backend has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (device < 0) {
        // one process per GPU: torchrun exports LOCAL_RANK; honour a device the host already selected
        const char* lr = getenv("LOCAL_RANK");
        int cur = 0;
        if (lr && *lr) device = atoi(lr) % count;
        else if (cudaGetDevice(&cur) == cudaSuccess) device = cur;
        else device = 0;
    }
    if (device >= count) return fail(TRN_GPU_ERROR, "device %d out of range (%d visible)", device, count);
    TRN_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    TRN_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10)
        return fail(TRN_UNSUPPORTED_BACKEND,
                    "Backend not supported on this platform: CUDA sm_%d%d (kernels are built for sm_100a only)",
                    prop.major, prop.minor);
    Context* c = new Context();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    c->cc_major = prop.major;
    c->cc_minor = prop.minor;
    c->hbm_bytes = prop.totalGlobalMem;
    snprintf(c->name, sizeof c->name, "%s", prop.name);
    TRN_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    TRN_CUDA(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
    TRN_CUDA(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    // keep freed scratch in the pool: GEMM operand splits are re-used call after call
    cudaMemPool_t pool;
    TRN_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t keep = UINT64_MAX;
    TRN_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    g_ctx = c;
    return TRN_OK;
}

// The current CUDA device is per-THREAD state: a worker thread that calls into the backend for the first time (the
// reference's rayon workers do, src/vector.rs:377-383) still has device 0 current, and its launches on the backend's
// streams would fail with "invalid resource handle" in a process bound to another GPU.  Every entry point goes through
// here, so this is where the calling thread is bound to the backend's device (one cudaGetDevice per call).
Context* ctx() {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx && init_locked(-1) != TRN_OK) return nullptr;
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != g_ctx->device) {
        if (cudaSetDevice(g_ctx->device) != cudaSuccess) {
            fail(TRN_GPU_ERROR, "cannot bind the calling thread to CUDA device %d", g_ctx->device);
            cudaGetLastError();
            return nullptr;
        }
    }
    return g_ctx;
}

std::mutex& host_mutex() {
    static std::mutex mu;
    return mu;
}

cudaStream_t resolve_stream(void* s) {
    if (s) return (cudaStream_t)s;
    Context* c = ctx();
    return c ? c->stream : nullptr;
}

Workspace* workspace(cudaStream_t s) {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g_ws.find(s);
    if (it != g_ws.end()) return it->second;
    Workspace* w = new Workspace();
    char* dev = nullptr;
    size_t bytes = kMaxReduceBlocks * (sizeof(float) + sizeof(uint64_t)) + 256;
    if (cudaMalloc(&dev, bytes) != cudaSuccess || cudaMemset(dev, 0, bytes) != cudaSuccess) {
        delete w;
        cudaGetLastError();
        return nullptr;
    }
    w->partial_idx = (uint64_t*)dev;
    w->partial_val = (float*)(dev + kMaxReduceBlocks * sizeof(uint64_t));
    char* tail = dev + kMaxReduceBlocks * (sizeof(float) + sizeof(uint64_t));
    w->ticket = (unsigned*)tail;
    w->scalar_u64 = (uint64_t*)(tail + 64);
    w->scalar_f32 = (float*)(tail + 128);
    char* host = nullptr;
    if (cudaMallocHost(&host, 128) != cudaSuccess) {
        cudaFree(dev);
        delete w;
        cudaGetLastError();
        return nullptr;
    }
    w->host_u64 = (uint64_t*)host;
    w->host_f32 = (float*)(host + 64);
    g_ws[s] = w;
    return w;
}

int scratch_alloc(void** p, size_t bytes, cudaStream_t s) {
    TRN_CUDA(cudaMallocAsync(p, bytes ? bytes : 16, s));
    return TRN_OK;
}
int scratch_free(void* p, cudaStream_t s) {
    if (p) TRN_CUDA(cudaFreeAsync(p, s));
    return TRN_OK;
}

// ---- transfers ------------------------------------------------------------------------------------
static bool is_pinned(const void* p);
bool is_pinned_host(const void* p) { return is_pinned(p); }
static bool is_pinned(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

static int ensure_stage() {
    if (g_stage[0]) return TRN_OK;
    for (int i = 0; i < 2; ++i) {
        TRN_CUDA(cudaMallocHost(&g_stage[i], kStageBytes));
        TRN_CUDA(cudaEventCreateWithFlags(&g_stage_ev[i], cudaEventDisableTiming));
    }
    return TRN_OK;
}

// Host -> HBM.  Pinned sources go straight to the copy engine; pageable sources are pipelined
// through two pinned 32 MiB stages so the CPU memcpy of chunk i+1 overlaps the DMA of chunk i.
int upload(float* dst, const float* src, size_t n, cudaStream_t s) {
    size_t bytes = n * sizeof(float);
    if (bytes == 0) return TRN_OK;
    if (is_pinned(src)) {
        TRN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
        return TRN_OK;
    }
    std::lock_guard<std::mutex> lk(g_stage_mu);
    TRN_TRY(ensure_stage());
    const char* sp = (const char*)src;
    char* dp = (char*)dst;
    int slot = 0;
    for (size_t off = 0; off < bytes; off += kStageBytes, slot ^= 1) {
        size_t len = bytes - off < kStageBytes ? bytes - off : kStageBytes;
        TRN_CUDA(cudaEventSynchronize(g_stage_ev[slot]));  // previous DMA out of this slot is done
        memcpy(g_stage[slot], sp + off, len);
        TRN_CUDA(cudaMemcpyAsync(dp + off, g_stage[slot], len, cudaMemcpyHostToDevice, s));
        TRN_CUDA(cudaEventRecord(g_stage_ev[slot], s));
    }
    return TRN_OK;
}

// HBM -> host.  Returns after the data is in `dst` (host-slice calls are synchronous).
int download(float* dst, const float* src, size_t n, cudaStream_t s) {
    size_t bytes = n * sizeof(float);
    if (bytes == 0) {
        TRN_CUDA(cudaStreamSynchronize(s));
        return TRN_OK;
    }
    if (is_pinned(dst)) {
        TRN_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
        TRN_CUDA(cudaStreamSynchronize(s));
        return TRN_OK;
    }
    std::lock_guard<std::mutex> lk(g_stage_mu);
    TRN_TRY(ensure_stage());
    char* dp = (char*)dst;
    const char* sp = (const char*)src;
    size_t nchunks = (bytes + kStageBytes - 1) / kStageBytes;
    auto chunk_len = [&](size_t i) { size_t off = i * kStageBytes; return bytes - off < kStageBytes ? bytes - off : kStageBytes; };
    // software pipeline: DMA of chunk i+1 overlaps the CPU memcpy of chunk i
    TRN_CUDA(cudaMemcpyAsync(g_stage[0], sp, chunk_len(0), cudaMemcpyDeviceToHost, s));
    TRN_CUDA(cudaEventRecord(g_stage_ev[0], s));
    for (size_t i = 0; i < nchunks; ++i) {
        int slot = (int)(i & 1);
        if (i + 1 < nchunks) {
            TRN_CUDA(cudaMemcpyAsync(g_stage[slot ^ 1], sp + (i + 1) * kStageBytes, chunk_len(i + 1),
                                     cudaMemcpyDeviceToHost, s));
            TRN_CUDA(cudaEventRecord(g_stage_ev[slot ^ 1], s));
        }
        TRN_CUDA(cudaEventSynchronize(g_stage_ev[slot]));
        memcpy(dp + i * kStageBytes, g_stage[slot], chunk_len(i));
    }
    TRN_CUDA(cudaStreamSynchronize(s));
    return TRN_OK;
}

}  // namespace trn

// =====================================================================================================
// C ABI: context, errors, buffers
// =====================================================================================================
using namespace trn;

struct trn_buf {
    float* dev;
    size_t len;
};

extern "C" {

int trn_cuda_init(int device) {
    std::lock_guard<std::mutex> lk(g_mu);
    return init_locked(device);
}

int trn_cuda_shutdown(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_ctx) return TRN_OK;
    cudaDeviceSynchronize();
    for (auto& kv : g_ws) {
        cudaFree(kv.second->partial_idx);
        cudaFreeHost(kv.second->host_u64);
        delete kv.second;
    }
    g_ws.clear();
    for (int i = 0; i < 2; ++i)
        if (g_stage[i]) {
            cudaFreeHost(g_stage[i]);
            cudaEventDestroy(g_stage_ev[i]);
            g_stage[i] = nullptr;
        }
    cudaStreamDestroy(g_ctx->stream);
    cudaStreamDestroy(g_ctx->copy_stream);
    cudaStreamDestroy(g_ctx->d2h_stream);
    delete g_ctx;
    g_ctx = nullptr;
    return TRN_OK;
}

int trn_cuda_is_available(void) { return ctx() != nullptr; }

int trn_device_count(int* count) {
    int c = 0;
    if (cudaGetDeviceCount(&c) != cudaSuccess) {
        cudaGetLastError();
        c = 0;
    }
    if (count) *count = c;
    return TRN_OK;
}

int trn_device_info(char* name, size_t cap, int* sm_count, uint64_t* hbm_bytes) {
    Context* c = ctx();
    if (!c) return TRN_GPU_ERROR;
    if (name && cap) snprintf(name, cap, "%s", c->name);
    if (sm_count) *sm_count = c->sm_count;
    if (hbm_bytes) *hbm_bytes = c->hbm_bytes;
    return TRN_OK;
}

size_t trn_last_error(char* buf, size_t cap) {
    size_t len = strlen(t_msg);
    if (buf && cap) {
        size_t ncopy = len < cap - 1 ? len : cap - 1;
        memcpy(buf, t_msg, ncopy);
        buf[ncopy] = 0;
    }
    return len;
}

void trn_last_mismatch(uint64_t* expected, uint64_t* actual) {
    if (expected) *expected = t_expected;
    if (actual) *actual = t_actual;
}

int trn_synchronize(void* stream) {
    if (!ctx()) return TRN_GPU_ERROR;
    TRN_CUDA(cudaStreamSynchronize(resolve_stream(stream)));
    return TRN_OK;
}

uint64_t trn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int trn_buf_alloc(size_t len, trn_buf** out) {
    if (!out) return fail(TRN_INVALID_INPUT, "trn_buf_alloc: null output handle");
    if (!ctx()) return TRN_GPU_ERROR;
    trn_buf* b = new trn_buf{nullptr, len};
    cudaError_t e = cudaMalloc(&b->dev, (len ? len : 1) * sizeof(float));
    if (e != cudaSuccess) {
        delete b;
        return fail_cuda(e, "cudaMalloc");
    }
    *out = b;
    return TRN_OK;
}

int trn_buf_free(trn_buf* buf) {
    if (!buf) return TRN_OK;
    cudaError_t e = cudaFree(buf->dev);
    delete buf;
    if (e != cudaSuccess) return fail_cuda(e, "cudaFree");
    return TRN_OK;
}

int trn_buf_upload(trn_buf* buf, const float* host, size_t len) {
    Context* c = ctx();
    if (!c) return TRN_GPU_ERROR;
    if (!buf) return fail(TRN_INVALID_INPUT, "trn_buf_upload: null buffer");
    if (len != buf->len) return fail_mismatch(buf->len, len);
    std::lock_guard<std::mutex> host_lock(host_mutex());
    TRN_TRY(upload(buf->dev, host, len, c->stream));
    TRN_CUDA(cudaStreamSynchronize(c->stream));  // the borrowed host slice may be dropped on return
    return TRN_OK;
}

int trn_buf_download(const trn_buf* buf, float* host, size_t len) {
    Context* c = ctx();
    if (!c) return TRN_GPU_ERROR;
    if (!buf) return fail(TRN_INVALID_INPUT, "trn_buf_download: null buffer");
    if (len != buf->len) return fail_mismatch(buf->len, len);
    std::lock_guard<std::mutex> host_lock(host_mutex());
    return download(host, buf->dev, len, c->stream);
}

size_t trn_buf_len(const trn_buf* buf) { return buf ? buf->len : 0; }
float* trn_buf_ptr(const trn_buf* buf) { return buf ? buf->dev : nullptr; }

int trn_host_alloc(size_t len, float** out) {
    if (!out) return fail(TRN_INVALID_INPUT, "trn_host_alloc: null output");
    if (!ctx()) return TRN_GPU_ERROR;
    TRN_CUDA(cudaMallocHost(out, (len ? len : 1) * sizeof(float)));
    return TRN_OK;
}

int trn_host_free(float* ptr) {
    if (ptr) TRN_CUDA(cudaFreeHost(ptr));
    return TRN_OK;
}

}  // extern "C"
