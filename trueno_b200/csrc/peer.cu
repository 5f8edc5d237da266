// peer.cu — fused "slice reduction + exchange" over NVLink peer memory.
//
// The sharded reductions (dot / sum / norm_l2 slices + all_reduce, arg slices + all_gather) end in a tiny
// exchange: one scalar (or one (value, index) pair) per rank.  Behind NCCL that exchange is a second launch and
// ~20 us of latency on top of a 75 us slice kernel at 8 GPUs.  Here the reduction kernel does it itself: its last
// block writes the slice result into every peer's mailbox with P2P stores and folds all ranks' results in rank
// order (reduce.cu: peer_exchange).  This file owns the mailboxes: one 2 KiB cudaMalloc per process, shared with
// the other ranks of the node through CUDA IPC handles that the host exchanges once (trueno_b200/parallel.py does
// it with one torch.distributed all_gather).
//
// Rules (as for any collective): every rank calls the same trn_*_allreduce / allgather entry points in the same
// order on ONE stream per process; world <= 8 (one NVSwitch domain).
#include <cstring>
#include <mutex>

#include "common.cuh"

using namespace trn;

struct trn_comm {
    int rank = 0, world = 1;
    unsigned long long* box[kMaxPeers] = {};
    bool opened[kMaxPeers] = {};
    unsigned seq = 0;
    std::mutex mu;
};

namespace {
constexpr size_t kMailboxBytes = 2 * kMaxPeers * 4 * sizeof(unsigned long long);   // parity x rank x word
unsigned long long* g_local_box = nullptr;   // this process's mailbox (one per process, reused by every comm)
std::mutex g_box_mu;

int ensure_local_box() {
    std::lock_guard<std::mutex> lk(g_box_mu);
    if (g_local_box) return TRN_OK;
    TRN_CUDA(cudaMalloc(&g_local_box, kMailboxBytes));
    TRN_CUDA(cudaMemset(g_local_box, 0, kMailboxBytes));
    TRN_CUDA(cudaDeviceSynchronize());
    return TRN_OK;
}

PeerCtx next_call(trn_comm* c) {
    std::lock_guard<std::mutex> lk(c->mu);
    PeerCtx pc = {};
    for (int r = 0; r < c->world; ++r) pc.box[r] = c->box[r];
    pc.rank = c->rank;
    pc.world = c->world;
    pc.seq = ++c->seq;
    if (pc.seq == 0) pc.seq = c->seq = 1;   // 0 is the "never written" state of a fresh mailbox
    return pc;
}
}  // namespace

extern "C" {

// Allocates this process's mailbox (once) and returns its 64-byte CUDA IPC handle for the peers.
int trn_comm_local_handle(void* handle64) {
    if (!handle64) return fail(TRN_INVALID_INPUT, "trn_comm_local_handle: null output");
    if (!ctx()) return TRN_GPU_ERROR;
    TRN_TRY(ensure_local_box());
    cudaIpcMemHandle_t h;
    TRN_CUDA(cudaIpcGetMemHandle(&h, g_local_box));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, 64);
    return TRN_OK;
}

// `handles`: world x 64 bytes, rank order (the all_gather of every rank's trn_comm_local_handle()).
int trn_comm_create(int rank, int world, const void* handles, trn_comm** out) {
    if (!out || !handles) return fail(TRN_INVALID_INPUT, "trn_comm_create: null argument");
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world)
        return fail(TRN_INVALID_INPUT, "trn_comm_create: rank %d / world %d outside 1..%d", rank, world, kMaxPeers);
    if (!ctx()) return TRN_GPU_ERROR;
    TRN_TRY(ensure_local_box());
    trn_comm* c = new trn_comm();
    c->rank = rank;
    c->world = world;
    for (int r = 0; r < world; ++r) {
        if (r == rank) { c->box[r] = g_local_box; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + 64 * r, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int q = 0; q < r; ++q) if (c->opened[q]) cudaIpcCloseMemHandle(c->box[q]);
            delete c;
            return fail_cuda(e, "cudaIpcOpenMemHandle (peer mailbox)");
        }
        c->box[r] = (unsigned long long*)p;
        c->opened[r] = true;
    }
    *out = c;
    return TRN_OK;
}

int trn_comm_destroy(trn_comm* c) {
    if (!c) return TRN_OK;
    if (ctx()) cudaDeviceSynchronize();
    for (int r = 0; r < c->world; ++r) if (c->opened[r]) cudaIpcCloseMemHandle(c->box[r]);
    delete c;
    return TRN_OK;
}

// ---- fused slice reduction + exchange: every rank receives the whole-vector result in `out` (device memory) ----
int trn_sum_allreduce_f32_dev(trn_comm* comm, const float* a, size_t n, float* out, void* stream) {
    if (!comm) return fail(TRN_INVALID_INPUT, "null communicator");
    if (!ctx()) return TRN_GPU_ERROR;
    const PeerCtx pc = next_call(comm);
    return launch_reduce(Reduce::Sum, a, nullptr, n, out, resolve_stream(stream), &pc);
}
int trn_dot_allreduce_f32_dev(trn_comm* comm, const float* a, size_t na, const float* b, size_t nb, float* out, void* stream) {
    if (!comm) return fail(TRN_INVALID_INPUT, "null communicator");
    if (na != nb) return fail_mismatch(na, nb);
    if (!ctx()) return TRN_GPU_ERROR;
    const PeerCtx pc = next_call(comm);
    return launch_reduce(Reduce::Dot, a, b, na, out, resolve_stream(stream), &pc);
}
// sqrt(sum over ALL slices of x^2): the exchange carries the sums of squares, the sqrt follows the fold
int trn_norm_l2_allreduce_f32_dev(trn_comm* comm, const float* a, size_t n, float* out, void* stream) {
    if (!comm) return fail(TRN_INVALID_INPUT, "null communicator");
    if (!ctx()) return TRN_GPU_ERROR;
    const PeerCtx pc = next_call(comm);
    return launch_reduce(Reduce::NormL2, a, nullptr, n, out, resolve_stream(stream), &pc);
}
// whole-vector argmax / argmin of a sharded vector: slice kernel + in-kernel exchange of (value, global index)
int trn_argmax_allgather_f32_dev(trn_comm* comm, const float* a, size_t n, uint64_t slice_start, uint64_t* out_idx,
                                 float* out_value, void* stream) {
    if (!comm) return fail(TRN_INVALID_INPUT, "null communicator");
    if (n == 0) return fail(TRN_INVALID_INPUT, "Empty vector");
    if (!ctx()) return TRN_GPU_ERROR;
    const PeerCtx pc = next_call(comm);
    return launch_argreduce(1, a, n, out_idx, out_value, resolve_stream(stream), slice_start == 0, slice_start, nullptr, &pc);
}
int trn_argmin_allgather_f32_dev(trn_comm* comm, const float* a, size_t n, uint64_t slice_start, uint64_t* out_idx,
                                 float* out_value, void* stream) {
    if (!comm) return fail(TRN_INVALID_INPUT, "null communicator");
    if (n == 0) return fail(TRN_INVALID_INPUT, "Empty vector");
    if (!ctx()) return TRN_GPU_ERROR;
    const PeerCtx pc = next_call(comm);
    return launch_argreduce(0, a, n, out_idx, out_value, resolve_stream(stream), slice_start == 0, slice_start, nullptr, &pc);
}

}  // extern "C"
