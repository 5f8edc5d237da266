"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed (NCCL over
NVLink 5 / NVSwitch) for the plumbing, the C-ABI kernels for the work.

The reference has no multi-GPU code at all; this module implements SURVEY.md §8e:

  op                                   partitioning                         exchange step
  -----------------------------------  -----------------------------------  ---------------------------------
  dot / sum / norm_l2                  contiguous slices of the vector      all_reduce(SUM) of one f32 partial
  min / max                            contiguous slices                    all_reduce(MIN/MAX) of one f32
  argmin / argmax                      contiguous slices, global u64 index  all_gather of (value, index) pairs,
                                                                            best value then LOWEST index wins
  softmax / log_softmax rows, maps     row / slice blocks                   none
  ONE softmax / log_softmax vector     contiguous slices of the vector      all_gather of (max, sum-of-exp) pairs,
                                                                            folded in rank order by every rank
  batched_matmul_4d                    contiguous ranges of batch*head      none
  matmul (large)                       C / A row blocks of whole 256-row    none on the product (one broadcast of B
                                       tiles, B replicated (ShardedMatrix)  when only one rank holds it)
  matvec                               A row blocks, v replicated           none (all_gather of y on request)

Partials are produced by the `_dev` kernels on the CURRENT torch stream and the collective is
enqueued on the same stream right behind them (no host synchronisation in between).  Tensors are
torch CUDA tensors used purely as device memory.  The combine functions accept CPU tensors too,
which is how the world_size-2 gloo tests exercise the exchange logic without a GPU; nothing here
ever computes a hot-path op on the CPU.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass(frozen=True)
class Shard:
    """Contiguous share [start, start + count) of `total` units owned by `rank` of `world`."""
    start: int
    count: int
    total: int
    rank: int
    world: int


def shard_range(total: int, rank: int, world: int, align: int = 1) -> Shard:
    """Even contiguous partition; every shard start is a multiple of `align` (128-bit loads want
    align=4 for f32 slices).  The first `rem` ranks get one extra aligned block."""
    blocks = (total + align - 1) // align
    base, rem = divmod(blocks, world)
    b0 = rank * base + min(rank, rem)
    nb = base + (1 if rank < rem else 0)
    start = min(b0 * align, total)
    end = min((b0 + nb) * align, total)
    return Shard(start, end - start, total, rank, world)


def init_distributed(backend: str | None = None) -> tuple[int, int, int]:
    """Reads RANK / LOCAL_RANK / WORLD_SIZE / MASTER_* (torchrun) and joins the process group.
    Returns (rank, local_rank, world).  world == 1 without env vars: no process group."""
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29531")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend, rank=rank, world_size=world,
                                    device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, local_rank, world


def current_stream_handle() -> int:
    """cudaStream_t of torch's current stream for the C ABI.  torch reports its default stream as 0,
    but 0 / NULL means "the backend's own stream" to the C ABI, so the legacy default stream is
    passed as cudaStreamLegacy (0x1)."""
    return torch.cuda.current_stream().cuda_stream or 1


def world_size() -> int:
    return dist.get_world_size() if dist.is_initialized() else 1


# ---- exchange steps (device-agnostic: CUDA tensors over NCCL, CPU tensors over gloo) ------------
def combine_sum(partial: torch.Tensor) -> torch.Tensor:
    """all_reduce(SUM) of per-slice partials (dot, sum, sum of squares), in place."""
    if world_size() > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.SUM)
    return partial


def combine_extreme(partial: torch.Tensor, is_max: bool) -> torch.Tensor:
    if world_size() > 1:
        dist.all_reduce(partial, op=dist.ReduceOp.MAX if is_max else dist.ReduceOp.MIN)
    return partial


NO_CANDIDATE = torch.iinfo(torch.int64).max


def combine_arg(value: torch.Tensor, index: torch.Tensor, is_max: bool) -> tuple[torch.Tensor, torch.Tensor]:
    """(value, GLOBAL index) pairs from every slice -> the scalar-backend answer for the whole vector
    (src/backends/scalar.rs:140-166).  Slice 0 carries the a[0] seed rule, so a NaN value from it is
    final (a NaN seed never loses).  Interior slices report NO_CANDIDATE when they hold nothing but
    NaNs / identity values (a NaN element never wins).  Otherwise: best value, and among equal
    values the LOWEST global index.  NCCL has no arg-reduce: this is an all_gather of `world` pairs.
    Tensor-op statement of the rule, used over gloo on CPU; on GPUs ShardedVector._arg runs the same rule as
    one kernel (trn_arg_combine_f32_dev) behind a single packed all_gather."""
    w = world_size()
    if w == 1:
        return value, index
    vals = torch.empty(w, dtype=value.dtype, device=value.device)
    idxs = torch.empty(w, dtype=index.dtype, device=index.device)
    dist.all_gather_into_tensor(vals, value.reshape(1))
    dist.all_gather_into_tensor(idxs, index.reshape(1))
    return pick_arg(vals, idxs, is_max)


def pick_arg(vals: torch.Tensor, idxs: torch.Tensor, is_max: bool) -> tuple[torch.Tensor, torch.Tensor]:
    """The selection rule of combine_arg on already-gathered pairs (pure tensor ops, no host sync)."""
    nan0 = torch.isnan(vals[0])
    usable = (idxs != NO_CANDIDATE) & ~torch.isnan(vals)
    worst = -float("inf") if is_max else float("inf")
    key = torch.where(usable, vals, torch.full_like(vals, worst))
    best = key.max() if is_max else key.min()
    cand = torch.where(usable & (key == best), idxs, torch.full_like(idxs, NO_CANDIDATE))
    out_v = torch.where(nan0, vals[0], best)
    out_i = torch.where(nan0, idxs[0], cand.min())
    return out_v.reshape(1), out_i.reshape(1)


# ---- fused exchange over NVLink peer memory ------------------------------------------------------------
def gather_softmax_pairs(pair: torch.Tensor) -> torch.Tensor:
    """all_gather of one (max, sum of exp(x - max)) pair per rank -> [world, 2] in rank order (CPU tensors under gloo
    in the tests; on the GPU the collective is enqueued on the current stream behind the stats kernel)."""
    w = world_size()
    if w == 1:
        return pair.reshape(1, 2)
    out = torch.empty(2 * w, dtype=pair.dtype, device=pair.device)
    dist.all_gather_into_tensor(out, pair.reshape(2).contiguous())
    return out.reshape(w, 2)


class PeerComm:
    """Peer mailboxes for the fused "slice reduction + exchange" kernels (csrc/peer.cu, csrc/reduce.cu).
    Every rank exports a 64-byte CUDA IPC handle; ONE all_gather of those handles at construction is all the
    collective plumbing the fused path needs — afterwards dot / sum / norm_l2 / argmax / argmin of a sharded
    vector are a single kernel launch per rank (P2P stores over NVLink, no NCCL call per reduction)."""

    def __init__(self):
        import ctypes as C

        import trueno_b200 as trn
        if not (dist.is_initialized() and dist.get_world_size() > 1):
            raise RuntimeError("PeerComm needs an initialised process group with world_size > 1")
        rank, world = dist.get_rank(), dist.get_world_size()
        self._h = None
        on_gpu = dist.get_backend() == "nccl"
        # Every rank takes part in both collectives below whatever happened locally, and all ranks fail TOGETHER: a rank
        # that cannot export or map a mailbox (no P2P path to a peer) must not leave the others waiting in a collective.
        buf = (C.c_ubyte * 64)()
        err = None
        try:
            trn.check(trn.lib.trn_comm_local_handle(buf))
        except trn.TruenoError as e:
            err = e
        mine = torch.tensor(list(buf) + [0 if err is None else 1], dtype=torch.uint8)
        if on_gpu:
            mine = mine.cuda()
        gathered = torch.empty(65 * world, dtype=torch.uint8, device=mine.device)
        dist.all_gather_into_tensor(gathered, mine)
        table = gathered.cpu().numpy().reshape(world, 65)
        h = C.c_void_p()
        if err is None and not table[:, 64].any():
            try:
                trn.check(trn.lib.trn_comm_create(rank, world, bytes(table[:, :64].tobytes()), C.byref(h)))
            except trn.TruenoError as e:
                err = e
        elif err is None:
            err = RuntimeError(f"rank {int(table[:, 64].argmax())} could not export its peer mailbox")
        failed = torch.tensor([0 if err is None else 1], dtype=torch.int32, device=mine.device)
        dist.all_reduce(failed, op=dist.ReduceOp.MAX)   # also: every rank has mapped every mailbox before the first fused call
        if int(failed.item()):
            if h.value:
                trn.lib.trn_comm_destroy(h)
            raise err if err is not None else RuntimeError("a peer rank could not map the peer mailboxes")
        self._h, self.rank, self.world = h, rank, world

    @classmethod
    def try_create(cls) -> "PeerComm | None":
        """A PeerComm, or None on EVERY rank when the fused path is not available (some rank has no P2P path to a peer):
        callers then pass `comm=None` and the sharded reductions exchange through NCCL."""
        try:
            return cls()
        except Exception:
            return None

    @property
    def handle(self):
        return self._h

    def close(self):
        import trueno_b200 as trn
        if self._h is not None:
            trn.lib.trn_comm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- sharded ops on device-resident slices --------------------------------------------------------
class ShardedVector:
    """This rank's contiguous slice of a global f32 vector, resident in HBM."""

    def __init__(self, local: torch.Tensor, shard: Shard, comm: "PeerComm | None" = None):
        """`comm`: use the fused peer-memory exchange (one launch per reduction) instead of kernel + NCCL."""
        assert local.is_cuda and local.dtype == torch.float32 and local.is_contiguous()
        self.local, self.shard, self.comm = local, shard, comm
        self._f32 = torch.zeros(1, dtype=torch.float32, device=local.device)
        self._i64 = torch.zeros(1, dtype=torch.int64, device=local.device)
        self._pair = torch.zeros(2, dtype=torch.int64, device=local.device)   # one trn_arg_pair (16 bytes)
        self._gathered = None

    def _stream(self) -> int:
        return current_stream_handle()

    def _partial(self, fn_name: str, other: "ShardedVector | None" = None) -> torch.Tensor:
        import trueno_b200 as trn
        fn = getattr(trn.lib, fn_name)
        n = self.local.numel()
        if other is None:
            trn.check(fn(self.local.data_ptr(), n, self._f32.data_ptr(), self._stream()))
        else:
            trn.check(fn(self.local.data_ptr(), n, other.local.data_ptr(), other.local.numel(),
                         self._f32.data_ptr(), self._stream()))
        return self._f32

    def _fused(self, fn_name: str, other: "ShardedVector | None" = None) -> torch.Tensor:
        import trueno_b200 as trn
        fn = getattr(trn.lib, fn_name)
        n = self.local.numel()
        if other is None:
            trn.check(fn(self.comm.handle, self.local.data_ptr(), n, self._f32.data_ptr(), self._stream()))
        else:
            trn.check(fn(self.comm.handle, self.local.data_ptr(), n, other.local.data_ptr(), other.local.numel(),
                         self._f32.data_ptr(), self._stream()))
        return self._f32

    def sum(self) -> torch.Tensor:
        if self.comm is not None:
            return self._fused("trn_sum_allreduce_f32_dev")
        return combine_sum(self._partial("trn_sum_f32_dev"))

    def dot(self, other: "ShardedVector") -> torch.Tensor:
        if self.comm is not None:
            return self._fused("trn_dot_allreduce_f32_dev", other)
        return combine_sum(self._partial("trn_dot_f32_dev", other))

    def norm_l2(self) -> torch.Tensor:
        if self.comm is not None:
            return self._fused("trn_norm_l2_allreduce_f32_dev")
        if world_size() == 1:
            return self._partial("trn_norm_l2_f32_dev")      # one launch, sqrt inside the kernel
        return combine_sum(self._partial("trn_sumsq_f32_dev")).sqrt_()

    def _arg(self, is_max: bool) -> tuple[torch.Tensor, torch.Tensor]:
        """slice kernel -> ONE all_gather of 16-byte (value, global index) pairs -> ONE combine kernel; all three are
        enqueued on the current stream with no host synchronisation (NCCL has no arg-reduce).
        An EMPTY vector is the reference's InvalidInput("Empty vector") on every rank alike (src/vector.rs:750-752); an
        empty SLICE (more ranks than aligned blocks) takes part in the exchange with "no candidate"."""
        import trueno_b200 as trn
        L = trn.lib
        w = world_size()
        if self.shard.total == 0:
            raise trn.TruenoError.InvalidInput("Empty vector")
        if self.comm is not None:   # fused: the slice kernel exchanges the pairs itself and applies the rule
            fused = L.trn_argmax_allgather_f32_dev if is_max else L.trn_argmin_allgather_f32_dev
            trn.check(fused(self.comm.handle, self.local.data_ptr(), self.local.numel(), self.shard.start,
                            self._i64.data_ptr(), self._f32.data_ptr(), self._stream()))
            return self._f32, self._i64
        if w == 1:   # the whole vector is here: one launch, no combine step
            one = L.trn_argmax_f32_dev if is_max else L.trn_argmin_f32_dev
            trn.check(one(self.local.data_ptr(), self.local.numel(), self._i64.data_ptr(), self._f32.data_ptr(), self._stream()))
            return self._f32, self._i64
        fn = L.trn_argmax_slice_pair_f32_dev if is_max else L.trn_argmin_slice_pair_f32_dev
        trn.check(fn(self.local.data_ptr(), self.local.numel(), self.shard.start, self._pair.data_ptr(), self._stream()))
        if w > 1:
            if self._gathered is None or self._gathered.numel() != 2 * w:
                self._gathered = torch.empty(2 * w, dtype=torch.int64, device=self.local.device)
            dist.all_gather_into_tensor(self._gathered, self._pair)
            pairs = self._gathered
        else:
            pairs = self._pair
        trn.check(L.trn_arg_combine_f32_dev(pairs.data_ptr(), w, int(is_max), self._i64.data_ptr(), self._f32.data_ptr(),
                                            self._stream()))
        return self._f32, self._i64

    def softmax(self, log: bool = False, out: "torch.Tensor | None" = None) -> torch.Tensor:
        """Vector::softmax / log_softmax (src/vector.rs:1516 / :1581) of the WHOLE sharded vector; returns this rank's
        slice of the result.  slice stats -> one all_gather of 8-byte pairs -> fold + write kernel, all on the current
        stream; every rank folds the same pairs in the same order, so all ranks normalise by identical bits."""
        import trueno_b200 as trn
        L = trn.lib
        if self.shard.total == 0:
            raise trn.TruenoError("EmptyVector")
        if out is None:
            out = torch.empty_like(self.local)
        n = self.local.numel()
        pair = torch.empty(2, dtype=torch.float32, device=self.local.device)
        trn.check(L.trn_softmax_slice_stats_f32_dev(self.local.data_ptr(), n, pair.data_ptr(), self._stream()))
        pairs = gather_softmax_pairs(pair)
        trn.check(L.trn_softmax_slice_apply_f32_dev(self.local.data_ptr(), n, pairs.data_ptr(), pairs.shape[0], int(log),
                                                    out.data_ptr(), self._stream()))
        return out

    def log_softmax(self, out: "torch.Tensor | None" = None) -> torch.Tensor:
        return self.softmax(log=True, out=out)

    def argmax(self) -> torch.Tensor:
        return self._arg(True)[1]

    def argmin(self) -> torch.Tensor:
        return self._arg(False)[1]

    def max(self) -> torch.Tensor:
        return self._arg(True)[0]

    def min(self) -> torch.Tensor:
        return self._arg(False)[0]


# ---- row-block sharded matrices (matmul by C row blocks, matvec by row blocks) -----------------------------
ROW_BLOCK = 256   # the reference's rayon unit (256-row blocks, src/matrix.rs:962-1011) == one tcgen05 CTA-pair tile of C


class ReplicatedMatrix:
    """A k x n f32 matrix every rank holds in full — the right-hand side of a row-block sharded product.  `prepare()`
    keeps its tf32 (hi, lo) split in HBM (trn_gemm_prepare_b_dev), so the products of all the row blocks — and every
    later product against it — split only their own rows of A."""

    def __init__(self, data: torch.Tensor, rows: int, cols: int):
        assert data.is_cuda and data.dtype == torch.float32 and data.is_contiguous() and data.numel() == rows * cols
        self.data, self.rows, self.cols = data, rows, cols
        self._handle = None

    @classmethod
    def broadcast(cls, data: "torch.Tensor | None", rows: int, cols: int, src: int = 0, device=None) -> "ReplicatedMatrix":
        """One broadcast of B from `src` to every rank (the only large transfer of the path)."""
        if data is None:
            data = torch.empty(rows * cols, dtype=torch.float32, device=device)
        if world_size() > 1:
            dist.broadcast(data, src=src)
        return cls(data.reshape(-1), rows, cols)

    def prepare(self) -> "ReplicatedMatrix":
        import ctypes as C

        import trueno_b200 as trn
        if self._handle is None:
            h = C.c_void_p()
            trn.check(trn.lib.trn_gemm_prepare_b_dev(self.data.data_ptr(), self.rows, self.cols, C.byref(h), current_stream_handle()))
            self._handle = h
        return self

    def close(self):
        import trueno_b200 as trn
        if self._handle is not None:
            trn.lib.trn_gemm_b_free(self._handle)
            self._handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class ShardedMatrix:
    """This rank's block of consecutive rows of a global rows x cols f32 matrix (row-major), resident in HBM.
    Row blocks are whole multiples of ROW_BLOCK rows (the last rank takes the ragged tail)."""

    def __init__(self, local: torch.Tensor, shard: Shard, cols: int):
        assert local.is_cuda and local.dtype == torch.float32 and local.is_contiguous() and local.numel() == shard.count * cols
        self.local, self.shard, self.cols = local, shard, cols

    @staticmethod
    def row_shard(rows: int, rank: "int | None" = None, world: "int | None" = None) -> Shard:
        if world is None:
            world = world_size()
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
        return shard_range(rows, rank, world, align=ROW_BLOCK)

    @property
    def rows(self) -> int:
        return self.shard.total

    def matmul(self, b: "ReplicatedMatrix", out: "torch.Tensor | None" = None) -> "ShardedMatrix":
        """Matrix::matmul (src/matrix.rs:285) with C sharded like A: every rank multiplies its row block by the whole
        of B — no collective, no partial sums, so every element of C is computed exactly as on one GPU (the kernel the
        WHOLE product would take, same k order: bit-identical to the unsharded product whatever the block heights).  Error text as the reference's (src/matrix.rs:286-291)."""
        import trueno_b200 as trn
        L = trn.lib
        if self.cols != b.rows:
            raise trn.TruenoError.InvalidInput(
                f"Matrix dimension mismatch for multiplication: {self.rows}\u00d7{self.cols} \u00d7 {b.rows}\u00d7{b.cols} "
                f"(inner dimensions {self.cols} and {b.rows} must match)")
        m = self.shard.count
        if out is None:
            out = torch.empty(m * b.cols, dtype=torch.float32, device=self.local.device)
        if m > 0 and b.cols > 0:
            # the row-block entry points pick the kernel by the WHOLE product's shape, so a short last block (or a one-row
            # block) does not fall to another kernel and another rounding
            if b._handle is not None:
                trn.check(L.trn_matmul_rowblock_prepared_f32_dev(self.local.data_ptr(), m, self.rows, self.cols, b._handle,
                                                                 out.data_ptr(), current_stream_handle()))
            else:
                trn.check(L.trn_matmul_rowblock_f32_dev(self.local.data_ptr(), m, self.rows, self.cols, b.data.data_ptr(), b.rows,
                                                        b.cols, out.data_ptr(), current_stream_handle()))
        return ShardedMatrix(out, self.shard, b.cols)

    def matvec(self, v: torch.Tensor, out: "torch.Tensor | None" = None) -> "ShardedVector":
        """Matrix::matvec (src/matrix.rs:1657; its rayon path splits the rows, :1676-1716): y = A v with v replicated and y
        sharded like the rows of A — no collective.  Returns the ShardedVector of this rank's rows of y."""
        import trueno_b200 as trn
        if v.numel() != self.cols:
            raise trn.TruenoError.InvalidInput(
                f"Vector length {v.numel()} does not match matrix columns {self.cols} for matrix-vector multiplication")
        m = self.shard.count
        if out is None:
            out = torch.empty(m, dtype=torch.float32, device=self.local.device)
        if m > 0:
            trn.check(trn.lib.trn_matvec_f32_dev(self.local.data_ptr(), m, self.cols, v.data_ptr(), v.numel(), out.data_ptr(),
                                                 current_stream_handle()))
        return ShardedVector(out, self.shard)

    def gather(self) -> torch.Tensor:
        """all_gather of the row blocks (the optional last step of section 8e): the whole matrix on every rank."""
        w = world_size()
        if w == 1:
            return self.local.reshape(self.shard.count, self.cols)
        counts = [self.row_shard(self.rows, r, w).count * self.cols for r in range(w)]
        return _gather_ragged(self.local.reshape(-1), counts).reshape(self.rows, self.cols)


def _gather_ragged(local: torch.Tensor, counts: "list[int]") -> torch.Tensor:
    """all_gather of slices of different (possibly zero) lengths: every rank pads to the longest slice, ONE
    all_gather_into_tensor, the padding is dropped on arrival."""
    w, longest = len(counts), max(max(counts), 1)
    padded = torch.zeros(longest, dtype=local.dtype, device=local.device)
    padded[:local.numel()] = local
    everything = torch.empty(w * longest, dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(everything, padded)
    return torch.cat([everything[r * longest:r * longest + counts[r]] for r in range(w)])


def gather_vector(v: "ShardedVector") -> torch.Tensor:
    """all_gather of the slices of a sharded vector (ragged and empty slices allowed; the partition is the one the
    vector was sharded with: `align` elements per block)."""
    w = world_size()
    if w == 1:
        return v.local
    sizes = torch.zeros(w, dtype=torch.int64, device=v.local.device)
    sizes[dist.get_rank()] = v.local.numel()
    dist.all_reduce(sizes)
    return _gather_ragged(v.local, [int(n) for n in sizes.tolist()])


# ---- CUDA graphs for launch-bound inner loops --------------------------------------------------------------------------
class CapturedLoop:
    """`iters` back-to-back calls of `fn` (C-ABI `_dev` launches, fused peer exchanges and NCCL collectives on the current
    stream) captured ONCE into a CUDA graph; replay() is one graph launch.  At 8 GPUs a sharded map or row kernel runs for
    ~20 us — shorter than the host takes to issue the next call — so the per-rank launch sequence goes into a graph
    (the Blackwell-native replacement for a tracing compiler: streams and graphs).  `fn` must not allocate or synchronise."""

    def __init__(self, fn, iters: int):
        self.iters = iters
        fn()                                   # warm-up outside capture: lazy workspaces, NCCL channels
        torch.cuda.current_stream().synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph, stream=torch.cuda.current_stream(), capture_error_mode="relaxed"):
            for _ in range(iters):
                fn()

    def replay(self):
        self.graph.replay()
