// peer.cu — fused "slice reduction + exchange" over NVLink peer memory.
//
// The sharded reductions (dot / sum / norm_l2 slices + all_reduce, arg slices + all_gather) end in a tiny
// exchange: one scalar (or one (value, index) pair) per rank.  Behind NCCL that exchange is a second launch and
// ~20 us of latency on top of a 75 us slice kernel at 8 GPUs.  Here the reduction kernel does it itself: its last
// block writes the slice result into every peer's mailbox with P2P stores and folds all ranks' results in rank
// order (reduce.cu: peer_exchange).  This file owns the mailboxes: ONE per communicator (trn_comm_local_handle
// allocates it, the trn_comm_create that follows adopts it), shared with the other ranks of the node through CUDA IPC
// handles that the host exchanges once (trueno_b200/parallel.py does it with one torch.distributed all_gather).
// The call sequence number lives in device memory beside the mailbox and is advanced by the kernel, so a chain of
// fused reductions can be captured in a CUDA graph and replayed.
//
// Rules (as for any collective): every rank calls the same trn_*_allreduce / allgather entry points in the same
// order on ONE stream per communicator; world <= 8 (one NVSwitch domain).  A peer that never shows up is an error,
// not a hang: an exchange gives up after TRN_PEER_TIMEOUT_MS (default 30 s), the result is NaN, and the communicator
// is poisoned — trn_comm_status and every later call on it return TRN_GPU_ERROR.
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <vector>

#include "common.cuh"

using namespace trn;

namespace {
constexpr size_t kSlotBytes = 2 * kMaxPeers * 4 * sizeof(unsigned long long);   // parity x rank x word
constexpr size_t kSeqOffset = 1024;                  // the device-side call counter, beside the slots
static_assert(kSlotBytes <= kSeqOffset, "the call counter must lie outside the message slots");
constexpr size_t kMailboxBytes = (size_t)2 << 20;    // one allocation granule, so every mailbox is its own IPC allocation

struct LocalBox {
    unsigned long long* dev = nullptr;
    cudaIpcMemHandle_t handle;
    bool adopted = false;   // owned by a communicator (freed with it); otherwise waiting for trn_comm_create
};
std::mutex g_box_mu;
std::vector<LocalBox> g_boxes;   // every mailbox this process has allocated and not yet freed
}  // namespace

struct trn_comm {
    int rank = 0, world = 1;
    unsigned long long* box[kMaxPeers] = {};
    bool opened[kMaxPeers] = {};
    unsigned* err_host = nullptr;   // pinned, mapped: {call number, rank + 1} of the first exchange that timed out
    unsigned* err_dev = nullptr;
    unsigned long long timeout_ns = 0;
};

namespace {
int check_comm(trn_comm* c) {
    if (!c) return fail(TRN_INVALID_INPUT, "null communicator");
    const volatile unsigned* e = c->err_host;
    if (e[0] != 0)
        return fail(TRN_GPU_ERROR, "peer exchange timed out waiting for rank %u in collective call %u: a peer is down or the "
                    "ranks' call sequences differ (the communicator is unusable)", e[1] ? e[1] - 1 : 0u, e[0]);
    return TRN_OK;
}
PeerCtx peer_ctx(const trn_comm* c) {
    PeerCtx pc = {};
    for (int r = 0; r < c->world; ++r) pc.box[r] = c->box[r];
    pc.seq = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(c->box[c->rank]) + kSeqOffset);
    pc.err = c->err_dev;
    pc.timeout_ns = c->timeout_ns;
    pc.rank = c->rank;
    pc.world = c->world;
    return pc;
}
}  // namespace

extern "C" {

// Allocates the mailbox of the NEXT communicator this process creates and returns its 64-byte CUDA IPC handle.
int trn_comm_local_handle(void* handle64) {
    if (!handle64) return fail(TRN_INVALID_INPUT, "trn_comm_local_handle: null output");
    if (!ctx()) return TRN_GPU_ERROR;
    LocalBox b;
    TRN_CUDA(cudaMalloc(&b.dev, kMailboxBytes));
    cudaError_t e = cudaMemset(b.dev, 0, kMailboxBytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&b.handle, b.dev);
    if (e != cudaSuccess) {
        cudaFree(b.dev);
        return fail_cuda(e, "peer mailbox allocation");
    }
    static_assert(sizeof(b.handle) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &b.handle, 64);
    std::lock_guard<std::mutex> lk(g_box_mu);
    g_boxes.push_back(b);
    return TRN_OK;
}

// `handles`: world x 64 bytes, rank order (the all_gather of every rank's trn_comm_local_handle()).  handles[rank] must
// be a handle this process obtained and has not used yet; a peer handle that names a mailbox of THIS process (several
// ranks driven from one process) is mapped directly instead of through IPC.
int trn_comm_create(int rank, int world, const void* handles, trn_comm** out) {
    if (!out || !handles) return fail(TRN_INVALID_INPUT, "trn_comm_create: null argument");
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
        return fail(TRN_INVALID_INPUT, "trn_comm_create: rank %d / world %d outside 1..%d", rank, world, kMaxPeers);
    if (!ctx()) return TRN_GPU_ERROR;
    trn_comm* c = new trn_comm();
    c->rank = rank;
    c->world = world;
    {
        const char* e = getenv("TRN_PEER_TIMEOUT_MS");
        const long ms = e ? atol(e) : 30000;
        c->timeout_ns = (unsigned long long)(ms > 0 ? ms : 30000) * 1000000ull;
    }
    auto local_box = [&](const char* h, bool adopt) -> unsigned long long* {
        std::lock_guard<std::mutex> lk(g_box_mu);
        for (LocalBox& b : g_boxes)
            if (memcmp(&b.handle, h, 64) == 0 && !(adopt && b.adopted)) {
                if (adopt) b.adopted = true;
                return b.dev;
            }
        return nullptr;
    };
    const char* hs = (const char*)handles;
    c->box[rank] = local_box(hs + 64 * rank, true);
    if (!c->box[rank]) {
        delete c;
        return fail(TRN_INVALID_INPUT, "trn_comm_create: handles[%d] is not an unused mailbox of this process "
                    "(call trn_comm_local_handle once per communicator)", rank);
    }
    cudaError_t e = cudaHostAlloc((void**)&c->err_host, 64, cudaHostAllocMapped);
    if (e == cudaSuccess) {
        memset(c->err_host, 0, 64);
        e = cudaHostGetDevicePointer((void**)&c->err_dev, c->err_host, 0);
    }
    for (int r = 0; r < world && e == cudaSuccess; ++r) {
        if (r == rank) continue;
        if (unsigned long long* same_process = local_box(hs + 64 * r, false)) { c->box[r] = same_process; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, hs + 64 * r, 64);
        void* p = nullptr;
        e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e == cudaSuccess) { c->box[r] = (unsigned long long*)p; c->opened[r] = true; }
    }
    if (e != cudaSuccess) {
        trn_comm_destroy(c);
        return fail_cuda(e, "trn_comm_create (peer mailbox mapping)");
    }
    *out = c;
    return TRN_OK;
}

int trn_comm_destroy(trn_comm* c) {
    if (!c) return TRN_OK;
    if (ctx()) cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r) if (c->opened[r]) cudaIpcCloseMemHandle(c->box[r]);
    if (c->err_host) cudaFreeHost(c->err_host);
    if (c->box[c->rank]) {
        std::lock_guard<std::mutex> lk(g_box_mu);
        for (size_t i = 0; i < g_boxes.size(); ++i)
            if (g_boxes[i].dev == c->box[c->rank]) {
                cudaFree(g_boxes[i].dev);
                g_boxes.erase(g_boxes.begin() + (long)i);
                break;
            }
    }
    delete c;
    return TRN_OK;
}

// TRN_OK, or TRN_GPU_ERROR once an exchange on this communicator has timed out (synchronise the stream first to learn
// about the calls still in flight).
int trn_comm_status(trn_comm* comm) { return check_comm(comm); }

// ---- fused slice reduction + exchange: every rank receives the whole-vector result in `out` (device memory) ----
// An EMPTY slice (n == 0: more ranks than aligned blocks) takes part with the identity, so no rank is left waiting.
int trn_sum_allreduce_f32_dev(trn_comm* comm, const float* a, size_t n, float* out, void* stream) {
    TRN_TRY(check_comm(comm));
    if (!ctx()) return TRN_GPU_ERROR;
    const PeerCtx pc = peer_ctx(comm);
    return launch_reduce(Reduce::Sum, a, nullptr, n, out, resolve_stream(stream), &pc);
}
int trn_dot_allreduce_f32_dev(trn_comm* comm, const float* a, size_t na, const float* b, size_t nb, float* out, void* stream) {
    TRN_TRY(check_comm(comm));
    if (na != nb) return fail_mismatch(na, nb);
    if (!ctx()) return TRN_GPU_ERROR;
    const PeerCtx pc = peer_ctx(comm);
    return launch_reduce(Reduce::Dot, a, b, na, out, resolve_stream(stream), &pc);
}
// sqrt(sum over ALL slices of x^2): the exchange carries the sums of squares, the sqrt follows the fold
int trn_norm_l2_allreduce_f32_dev(trn_comm* comm, const float* a, size_t n, float* out, void* stream) {
    TRN_TRY(check_comm(comm));
    if (!ctx()) return TRN_GPU_ERROR;
    const PeerCtx pc = peer_ctx(comm);
    return launch_reduce(Reduce::NormL2, a, nullptr, n, out, resolve_stream(stream), &pc);
}
// whole-vector argmax / argmin of a sharded vector: slice kernel + in-kernel exchange of (value, global index).
// An empty slice 0 means the whole vector is empty (slices are dealt front to back): InvalidInput("Empty vector") on
// every rank alike; an empty interior slice contributes "no candidate".
int trn_argmax_allgather_f32_dev(trn_comm* comm, const float* a, size_t n, uint64_t slice_start, uint64_t* out_idx,
                                 float* out_value, void* stream) {
    TRN_TRY(check_comm(comm));
    if (n == 0 && slice_start == 0) return fail(TRN_INVALID_INPUT, "Empty vector");
    if (!ctx()) return TRN_GPU_ERROR;
    const PeerCtx pc = peer_ctx(comm);
    return launch_argreduce(1, a, n, out_idx, out_value, resolve_stream(stream), slice_start == 0, slice_start, nullptr, &pc);
}
int trn_argmin_allgather_f32_dev(trn_comm* comm, const float* a, size_t n, uint64_t slice_start, uint64_t* out_idx,
                                 float* out_value, void* stream) {
    TRN_TRY(check_comm(comm));
    if (n == 0 && slice_start == 0) return fail(TRN_INVALID_INPUT, "Empty vector");
    if (!ctx()) return TRN_GPU_ERROR;
    const PeerCtx pc = peer_ctx(comm);
    return launch_argreduce(0, a, n, out_idx, out_value, resolve_stream(stream), slice_start == 0, slice_start, nullptr, &pc);
}

}  // extern "C"
