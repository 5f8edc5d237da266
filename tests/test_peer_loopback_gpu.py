"""The fused "slice reduction + exchange" protocol (csrc/peer.cu, reduce.cu peer_exchange; SURVEY.md §8e: slices + one
exchange step) on ONE GPU: both ranks of a world of 2 live in this process — two communicators, two mailboxes mapped
directly, two streams — so the driver's single-GPU box exercises the same kernels, mailbox slots, sequence numbers and
cross-slice rules that tests/dist_worker.py runs over NVLink.  Truths: f64 sums on the sum|terms| scale; argmax / argmin by
the scalar-backend rule (first occurrence, NaN seed wins; src/backends/scalar.rs:140-166).  Also here: a peer that never
shows up is TruenoError::GpuError after a bounded wait (never a hang), a second communicator starts from a clean mailbox,
and a chain of fused calls replays from a CUDA graph."""
import ctypes as C
import os
import subprocess
import sys
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")
f32 = np.float32
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pair_of_comms(trn):
    hs = [(C.c_ubyte * 64)(), (C.c_ubyte * 64)()]
    for h in hs:
        trn.check(trn.lib.trn_comm_local_handle(h))
    handles = bytes(hs[0]) + bytes(hs[1])
    comms = []
    for r in range(2):
        c = C.c_void_p()
        trn.check(trn.lib.trn_comm_create(r, 2, handles, C.byref(c)))
        comms.append(c)
    return comms


@pytest.fixture()
def world2(trn):
    trn.check(trn.lib.trn_cuda_init(0))
    torch.cuda.set_device(0)
    os.environ["TRN_PEER_TIMEOUT_MS"] = "8000"      # read by trn_comm_create: a protocol bug fails in seconds
    comms = _pair_of_comms(trn)
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    # Both "ranks" share one GPU and one host thread here: while rank 0's kernel waits for rank 1's message the host must be
    # free to launch rank 1's kernel, so nothing on that path may block on the device — the per-stream reduction
    # workspaces (cudaMalloc / cudaMallocHost on first use) are created up front, and every kernel the tests launch is
    # loaded up front too (CUDA loads a kernel lazily at its first launch, which waits for the running kernels).
    warm = torch.ones(64, device="cuda")
    L = trn.lib
    for s in streams:
        h = s.cuda_stream
        trn.check(L.trn_sum_f32_dev(warm.data_ptr(), 8, warm[8:].data_ptr(), h))
        trn.check(L.trn_dot_f32_dev(warm.data_ptr(), 8, warm.data_ptr(), 8, warm[8:].data_ptr(), h))
        trn.check(L.trn_norm_l2_f32_dev(warm.data_ptr(), 8, warm[8:].data_ptr(), h))
        trn.check(L.trn_argmax_f32_dev(warm.data_ptr(), 8, None, warm[8:].data_ptr(), h))
        trn.check(L.trn_argmin_f32_dev(warm.data_ptr(), 8, None, warm[8:].data_ptr(), h))
    torch.cuda.synchronize()
    yield comms, streams
    os.environ.pop("TRN_PEER_TIMEOUT_MS", None)
    torch.cuda.synchronize()
    for c in comms:
        trn.lib.trn_comm_destroy(c)


def _slices(trn, a, n):
    from trueno_b200.parallel import shard_range
    shards = [shard_range(n, r, 2, 4) for r in range(2)]
    return shards, [torch.from_numpy(a[s.start:s.start + s.count]).cuda() for s in shards]


@pytest.mark.parametrize("n", [1 << 22, (1 << 22) + 37, 1001, 6, 3])
def test_two_ranks_fused_reductions(trn, oracle, world2, n):
    from oracle import SCALAR
    comms, streams = world2
    L = trn.lib
    rng = np.random.default_rng(n)
    a = rng.uniform(-1, 1, n).astype(f32)
    b = rng.uniform(-1, 1, n).astype(f32)
    if n > 100:   # planted cross-rank ties: the lowest global index wins
        a[n - 5] = a[7] = f32(3.0)
        a[n - 9] = a[11] = f32(-4.0)
    shards, da = _slices(trn, a, n)
    _, db = _slices(trn, b, n)
    outs = [torch.zeros(4, device="cuda") for _ in range(2)]
    idxs = [torch.zeros(2, dtype=torch.int64, device="cuda") for _ in range(2)]
    tsum, asum = oracle.f64_sum(a)
    tdot, adot = oracle.f64_dot(a, b)
    for rep in range(3):    # repeated calls stay in step (call numbers, double-buffered slots)
        for r in range(2):
            s = streams[r].cuda_stream
            trn.check(L.trn_sum_allreduce_f32_dev(comms[r], da[r].data_ptr(), shards[r].count, outs[r][0:].data_ptr(), s))
            trn.check(L.trn_dot_allreduce_f32_dev(comms[r], da[r].data_ptr(), shards[r].count, db[r].data_ptr(), shards[r].count,
                                                  outs[r][1:].data_ptr(), s))
            trn.check(L.trn_norm_l2_allreduce_f32_dev(comms[r], da[r].data_ptr(), shards[r].count, outs[r][2:].data_ptr(), s))
            trn.check(L.trn_argmax_allgather_f32_dev(comms[r], da[r].data_ptr(), shards[r].count, shards[r].start, idxs[r][0:].data_ptr(),
                                                     outs[r][3:].data_ptr(), s))
            trn.check(L.trn_argmin_allgather_f32_dev(comms[r], da[r].data_ptr(), shards[r].count, shards[r].start, idxs[r][1:].data_ptr(),
                                                     None, s))
        torch.cuda.synchronize()
        assert torch.equal(outs[0], outs[1]) and torch.equal(idxs[0], idxs[1])      # same bits on both ranks
        o = outs[0].cpu().numpy()
        assert abs(float(o[0]) - tsum) <= 1e-5 * asum
        assert abs(float(o[1]) - tdot) <= 1e-5 * adot
        tn = float(np.sqrt(np.sum(a.astype(np.float64) ** 2)))
        assert abs(float(o[2]) - tn) <= 1e-5 * tn
        assert int(idxs[0][0]) == oracle.argmax(a, backend=SCALAR) and int(idxs[0][1]) == oracle.argmin(a, backend=SCALAR)
        assert float(o[3]) == float(a.max())
    for c in comms:
        trn.check(L.trn_comm_status(c))


def test_nan_seed_and_empty_slice_rules(trn, world2):
    comms, streams = world2
    L = trn.lib
    a = np.linspace(-1, 1, 4096).astype(f32)
    a[0] = np.nan                                     # a NaN seed (a[0]) never loses, on any rank
    shards, da = _slices(trn, a, a.size)
    idx = [torch.zeros(1, dtype=torch.int64, device="cuda") for _ in range(2)]
    for r in range(2):
        trn.check(L.trn_argmax_allgather_f32_dev(comms[r], da[r].data_ptr(), shards[r].count, shards[r].start, idx[r].data_ptr(), None,
                                                 streams[r].cuda_stream))
    torch.cuda.synchronize()
    assert int(idx[0]) == 0 and int(idx[1]) == 0
    # an empty slice 0 means an empty vector: the reference's error, before anything is launched
    st = L.trn_argmax_allgather_f32_dev(comms[0], None, 0, 0, idx[0].data_ptr(), None, streams[0].cuda_stream)
    assert st == 2 and trn.last_error() == "Empty vector"
    with pytest.raises(trn.TruenoError) as e:
        trn.check(st)
    assert e.value == trn.TruenoError.InvalidInput("Empty vector")


def test_second_communicator_starts_clean(trn, world2):
    """A communicator owns its mailbox: call numbers of an earlier communicator cannot satisfy a later one's wait."""
    comms, streams = world2
    L = trn.lib
    x = [torch.full((1024,), float(r + 1), device="cuda") for r in range(2)]
    out = [torch.zeros(1, device="cuda") for _ in range(2)]
    for _ in range(4):
        for r in range(2):
            trn.check(L.trn_sum_allreduce_f32_dev(comms[r], x[r].data_ptr(), 1024, out[r].data_ptr(), streams[r].cuda_stream))
    torch.cuda.synchronize()
    assert float(out[0]) == 3072.0
    second = _pair_of_comms(trn)
    y = [torch.full((1024,), float(10 * (r + 1)), device="cuda") for r in range(2)]
    for call in range(3):
        # rank 1 of the NEW pair is deliberately late: rank 0 must wait for it, not read a stale message
        trn.check(L.trn_sum_allreduce_f32_dev(second[0], y[0].data_ptr(), 1024, out[0].data_ptr(), streams[0].cuda_stream))
        time.sleep(0.02)                  # host-side delay: no kernel is loaded or launched while rank 0 waits
        trn.check(L.trn_sum_allreduce_f32_dev(second[1], y[1].data_ptr(), 1024, out[1].data_ptr(), streams[1].cuda_stream))
        torch.cuda.synchronize()
        assert float(out[0]) == float(out[1]) == 30720.0, call
    for c in second:
        trn.lib.trn_comm_destroy(c)


def test_fused_chain_replays_from_a_cuda_graph(trn, world2):
    """The call number lives in device memory, so captured fused reductions can be replayed (parallel.CapturedLoop)."""
    comms, streams = world2
    L = trn.lib
    x = [torch.arange(4096, device="cuda", dtype=torch.float32) * (r + 1) for r in range(2)]
    out = [torch.zeros(1, device="cuda") for _ in range(2)]
    graphs = []
    for r in range(2):
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=streams[r], capture_error_mode="relaxed"):
            for _ in range(5):
                trn.check(L.trn_sum_allreduce_f32_dev(comms[r], x[r].data_ptr(), 4096, out[r].data_ptr(), streams[r].cuda_stream))
        graphs.append(g)
    want = float(np.arange(4096, dtype=np.float64).sum() * 3)
    for _ in range(3):
        for r in range(2):
            with torch.cuda.stream(streams[r]):
                graphs[r].replay()
        torch.cuda.synchronize()
        assert float(out[0]) == float(out[1]) == want
    for c in comms:
        trn.check(L.trn_comm_status(c))


_TIMEOUT_WORKER = r"""
import ctypes as C, sys
sys.path.insert(0, %r)
import torch
import trueno_b200 as trn
trn.check(trn.lib.trn_cuda_init(0))
hs = [(C.c_ubyte * 64)(), (C.c_ubyte * 64)()]
for h in hs:
    trn.check(trn.lib.trn_comm_local_handle(h))
c = C.c_void_p()
trn.check(trn.lib.trn_comm_create(0, 2, bytes(hs[0]) + bytes(hs[1]), C.byref(c)))     # rank 1 never calls
x = torch.ones(4096, device="cuda"); out = torch.zeros(1, device="cuda")
trn.check(trn.lib.trn_sum_allreduce_f32_dev(c, x.data_ptr(), 4096, out.data_ptr(), None))
trn.check(trn.lib.trn_synchronize(None))                                               # returns: the wait is bounded
assert torch.isnan(out).all(), out
st = trn.lib.trn_comm_status(c)
assert st == 5 and "timed out waiting for rank 1" in trn.last_error(), (st, trn.last_error())
st = trn.lib.trn_sum_allreduce_f32_dev(c, x.data_ptr(), 4096, out.data_ptr(), None)   # poisoned: fails before launching
assert st == 5
try:
    trn.check(st)
except trn.TruenoError as e:
    assert e.variant == "GpuError"
print("timeout ok")
"""


def test_dead_peer_is_a_gpu_error_not_a_hang():
    env = dict(os.environ, TRN_PEER_TIMEOUT_MS="300", PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, "-c", _TIMEOUT_WORKER % ROOT], env=env, cwd=ROOT, stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=120)
    assert r.returncode == 0 and "timeout ok" in r.stdout, r.stdout[-3000:]


def test_calls_from_a_second_thread(trn):
    """The current CUDA device is per-thread state; every entry point binds the calling thread to the backend's device."""
    import threading
    trn.check(trn.lib.trn_cuda_init(0))
    res = {}

    def work():
        v = trn.Vector.from_slice(np.arange(1000, dtype=f32))
        res["sum"] = float(v.sum())
        res["mm"] = trn.Matrix.from_vec(2, 2, [1, 2, 3, 4]).matmul(trn.Matrix.from_vec(2, 2, [5, 6, 7, 8])).to_numpy().tolist()
    t = threading.Thread(target=work)
    t.start()
    t.join()
    assert res["sum"] == 499500.0 and res["mm"] == [[19.0, 22.0], [43.0, 50.0]]
