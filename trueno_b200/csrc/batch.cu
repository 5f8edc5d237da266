// batch.cu — device-resident op chaining: the CUDA counterpart of `GpuCommandBatch`
// (src/backends/gpu/batch.rs:54-1019: upload -> relu / scale / add / mul / dot / sigmoid / tanh / swish / gelu /
// sub -> execute -> read).  SURVEY.md 8f rank 1.
//
// The reference queues ops, then `execute()` creates one wgpu buffer per BufferId, uploads the inputs and
// submits one dispatch per op; `read()` maps a result back.  Here: every BufferId is a slice of ONE device
// arena; `execute()` uploads the inputs with pinned staging and launches the whole op sequence as ONE CUDA
// graph, captured from the same launchers the `_dev` entry points use (so results are bit-identical to calling
// them one by one) and instantiated once — re-executing a batch (after trn_batch_update) replays the graph
// with a single launch.  Nothing leaves HBM between ops; `read()` is the only synchronisation point.
#include <vector>

#include "common.cuh"

using namespace trn;

namespace {
enum BatchOp { B_RELU = 0, B_SCALE, B_ADD, B_MUL, B_DOT, B_SIGMOID, B_TANH, B_SWISH, B_GELU, B_SUB, B_COUNT };

struct BufInfo {
    size_t len = 0;
    size_t offset = 0;            // element offset into the arena (256-byte aligned)
    std::vector<float> host;      // uploads only: the copy `upload(&[f32])` takes (batch.rs:169-171)
    bool is_input = false;
};
struct OpInfo {
    int kind;
    uint32_t a, b, out;
    float scalar;
};
}  // namespace

struct trn_batch {
    std::vector<BufInfo> bufs;
    std::vector<OpInfo> ops;
    float* arena = nullptr;
    size_t arena_elems = 0;
    size_t planned_bufs = 0, planned_ops = 0;   // what the instantiated graph covers
    cudaGraphExec_t exec = nullptr;
    bool executed = false;
};

namespace {

int bad_id(const trn_batch* b, uint32_t id) {
    if (id >= b->bufs.size()) return fail(TRN_INVALID_INPUT, "Invalid buffer ID %u", id);
    return TRN_OK;
}

void drop_plan(trn_batch* b) {
    if (b->exec) { cudaGraphExecDestroy(b->exec); b->exec = nullptr; }
    if (b->arena) { cudaFree(b->arena); b->arena = nullptr; }
    b->arena_elems = 0;
    b->planned_bufs = b->planned_ops = 0;
    b->executed = false;
}

int enqueue_ops(trn_batch* b, cudaStream_t s) {
    for (const OpInfo& op : b->ops) {
        const float* pa = b->arena + b->bufs[op.a].offset;
        const float* pb = op.b != UINT32_MAX ? b->arena + b->bufs[op.b].offset : nullptr;
        float* po = b->arena + b->bufs[op.out].offset;
        const size_t n = b->bufs[op.a].len;
        switch (op.kind) {
            case B_RELU:    TRN_TRY(launch_map(Map::Relu, pa, nullptr, nullptr, po, n, 0.f, 0.f, s)); break;
            case B_SCALE:   TRN_TRY(launch_map(Map::Scale, pa, nullptr, nullptr, po, n, op.scalar, 0.f, s)); break;
            case B_ADD:     TRN_TRY(launch_map(Map::Add, pa, pb, nullptr, po, n, 0.f, 0.f, s)); break;
            case B_MUL:     TRN_TRY(launch_map(Map::Mul, pa, pb, nullptr, po, n, 0.f, 0.f, s)); break;
            case B_SUB:     TRN_TRY(launch_map(Map::Sub, pa, pb, nullptr, po, n, 0.f, 0.f, s)); break;
            case B_SIGMOID: TRN_TRY(launch_map(Map::Sigmoid, pa, nullptr, nullptr, po, n, 0.f, 0.f, s)); break;
            case B_TANH:    TRN_TRY(launch_map(Map::Tanh, pa, nullptr, nullptr, po, n, 0.f, 0.f, s)); break;
            case B_SWISH:   TRN_TRY(launch_map(Map::Swish, pa, nullptr, nullptr, po, n, 0.f, 0.f, s)); break;
            case B_GELU:    TRN_TRY(launch_map(Map::Gelu, pa, nullptr, nullptr, po, n, 0.f, 0.f, s)); break;
            case B_DOT:     TRN_TRY(launch_reduce(Reduce::Dot, pa, pb, n, po, s)); break;
            default: return fail(TRN_INVALID_INPUT, "unknown batch op %d", op.kind);
        }
    }
    return TRN_OK;
}

}  // namespace

extern "C" {

int trn_batch_create(trn_batch** out) {   // GpuCommandBatch::new (batch.rs:140)
    if (!out) return fail(TRN_INVALID_INPUT, "trn_batch_create: null output handle");
    if (!ctx()) return TRN_GPU_ERROR;
    *out = new trn_batch();
    return TRN_OK;
}

int trn_batch_destroy(trn_batch* b) {
    if (!b) return TRN_OK;
    if (ctx()) cudaStreamSynchronize(ctx()->stream);
    drop_plan(b);
    delete b;
    return TRN_OK;
}

int trn_batch_upload(trn_batch* b, const float* data, size_t len, uint32_t* id) {   // batch.rs:169
    if (!b || !id) return fail(TRN_INVALID_INPUT, "trn_batch_upload: null argument");
    BufInfo bi;
    bi.len = len;
    bi.host.assign(data, data + len);
    bi.is_input = true;
    b->bufs.push_back(std::move(bi));
    *id = (uint32_t)(b->bufs.size() - 1);
    return TRN_OK;
}

// Replaces the data of an uploaded buffer (same length) so an executed batch can be replayed on new inputs.
int trn_batch_update(trn_batch* b, uint32_t id, const float* data, size_t len) {
    if (!b) return fail(TRN_INVALID_INPUT, "trn_batch_update: null batch");
    TRN_TRY(bad_id(b, id));
    BufInfo& bi = b->bufs[id];
    if (!bi.is_input) return fail(TRN_INVALID_INPUT, "Buffer %u is not an uploaded buffer", id);
    if (len != bi.len) return fail_mismatch(bi.len, len);
    bi.host.assign(data, data + len);
    return TRN_OK;
}

// op: 0 relu, 1 scale, 2 add, 3 mul, 4 dot, 5 sigmoid, 6 tanh, 7 swish, 8 gelu, 9 sub (batch.rs:181-360).
// `b_id` is ignored by unary ops, `scalar` by everything but scale.  Binary ops on buffers of different
// sizes: the reference panics with "Buffer size mismatch: {} vs {}" (batch.rs:215-232); here InvalidInput.
int trn_batch_op(trn_batch* b, int op, uint32_t a_id, uint32_t b_id, float scalar, uint32_t* out_id) {
    if (!b || !out_id) return fail(TRN_INVALID_INPUT, "trn_batch_op: null argument");
    if (op < 0 || op >= B_COUNT) return fail(TRN_INVALID_INPUT, "unknown batch op %d", op);
    TRN_TRY(bad_id(b, a_id));
    const bool binary = op == B_ADD || op == B_MUL || op == B_DOT || op == B_SUB;
    if (binary) {
        TRN_TRY(bad_id(b, b_id));
        if (b->bufs[a_id].len != b->bufs[b_id].len)
            return fail(TRN_INVALID_INPUT, "Buffer size mismatch: %zu vs %zu", b->bufs[a_id].len, b->bufs[b_id].len);
    }
    BufInfo bo;
    bo.len = op == B_DOT ? 1 : b->bufs[a_id].len;   // dot returns a single-element buffer (batch.rs:275)
    b->bufs.push_back(std::move(bo));
    *out_id = (uint32_t)(b->bufs.size() - 1);
    b->ops.push_back(OpInfo{op, a_id, binary ? b_id : UINT32_MAX, *out_id, scalar});
    return TRN_OK;
}

size_t trn_batch_num_operations(const trn_batch* b) { return b ? b->ops.size() : 0; }   // batch.rs:1014
size_t trn_batch_num_buffers(const trn_batch* b) { return b ? b->bufs.size() : 0; }      // batch.rs:1019

int trn_batch_execute(trn_batch* b) {   // batch.rs:362
    if (!b) return fail(TRN_INVALID_INPUT, "trn_batch_execute: null batch");
    Context* c = ctx();
    if (!c) return TRN_GPU_ERROR;
    // the backend stream goes into capture below and carries the uploads: no host-slice call may touch it meanwhile
    std::lock_guard<std::mutex> host_lock(host_mutex());
    cudaStream_t s = c->stream;
    // (re)plan when buffers or ops were added since the last execute
    if (b->planned_bufs != b->bufs.size() || b->planned_ops != b->ops.size()) {
        TRN_CUDA(cudaStreamSynchronize(s));
        drop_plan(b);
        size_t off = 0;
        for (BufInfo& bi : b->bufs) {
            bi.offset = off;
            off += (bi.len + 63) & ~(size_t)63;
        }
        b->arena_elems = off ? off : 64;
        TRN_CUDA(cudaMalloc(&b->arena, b->arena_elems * sizeof(float)));
        if (!b->ops.empty()) {
            if (!workspace(s)) return fail(TRN_GPU_ERROR, "failed to allocate the reduction workspace");
            cudaGraph_t graph = nullptr;
            TRN_CUDA(cudaStreamBeginCapture(s, cudaStreamCaptureModeRelaxed));
            const uint64_t before = trn_launch_count();
            const int st = enqueue_ops(b, s);
            uncount_launch((unsigned)(trn_launch_count() - before));   // captured nodes are not launches yet
            const cudaError_t e = cudaStreamEndCapture(s, &graph);
            if (st != TRN_OK) { if (graph) cudaGraphDestroy(graph); return st; }
            TRN_CUDA(e);
            const cudaError_t ei = cudaGraphInstantiate(&b->exec, graph, 0);
            cudaGraphDestroy(graph);
            TRN_CUDA(ei);
        }
        b->planned_bufs = b->bufs.size();
        b->planned_ops = b->ops.size();
    }
    for (const BufInfo& bi : b->bufs)
        if (bi.is_input && bi.len) TRN_TRY(upload(b->arena + bi.offset, bi.host.data(), bi.len, s));
    if (b->exec) {
        TRN_CUDA(cudaGraphLaunch(b->exec, s));
        count_launch((unsigned)b->ops.size());
    }
    b->executed = true;
    return TRN_OK;
}

int trn_batch_read(trn_batch* b, uint32_t id, float* out, size_t len) {   // batch.rs:956
    if (!b) return fail(TRN_INVALID_INPUT, "trn_batch_read: null batch");
    TRN_TRY(bad_id(b, id));
    if (!b->executed || id >= b->planned_bufs) return fail(TRN_INVALID_INPUT, "Buffer not executed yet");
    if (len != b->bufs[id].len) return fail_mismatch(b->bufs[id].len, len);
    Context* c = ctx();
    if (!c) return TRN_GPU_ERROR;
    std::lock_guard<std::mutex> host_lock(host_mutex());
    return download(out, b->arena + b->bufs[id].offset, len, c->stream);
}

}  // extern "C"
