"""Host-side logic of the multi-GPU layer (trueno_b200/parallel.py) on CPU: the partitioner and
the cross-slice exchange steps over a world_size-2 (and 3) gloo group.  Per-slice partials are
produced here by the ORACLE (it stands in for the slice kernels, whose own parity is covered by
the -m gpu tests); what is under test is the combine rule and the collectives."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

f32 = np.float32


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_range_partitions_exactly():
    from trueno_b200.parallel import shard_range
    for total in (0, 1, 5, 255, 256, 1000, 1 << 20, (1 << 30)):
        for world in (1, 2, 3, 4, 8):
            for align in (1, 4):
                shards = [shard_range(total, r, world, align) for r in range(world)]
                assert shards[0].start == 0 and sum(s.count for s in shards) == total
                for a, b in zip(shards, shards[1:]):
                    assert a.start + a.count == b.start
                    assert b.start % align == 0 or b.count == 0
                assert max(s.count for s in shards) - min(s.count for s in shards) < 2 * align
    # BASELINE config 4: 2^30 over 8 GPUs -> 2^27 each; config 3: 256 heads -> 32 each
    assert [shard_range(1 << 30, r, 8, 4).count for r in range(8)] == [1 << 27] * 8
    assert [shard_range(256, r, 8).count for r in range(8)] == [32] * 8


def test_row_blocks_of_a_sharded_matrix_are_whole_tiles():
    """ShardedMatrix.row_shard: the reference's rayon unit is a 256-row block (src/matrix.rs:962-1011); every rank owns a
    run of whole blocks, the last non-empty one takes the ragged tail, nothing is lost or duplicated."""
    from trueno_b200.parallel import ROW_BLOCK, ShardedMatrix
    for rows in (0, 1, 255, 256, 1000, 1324, 8192, 32768, 32768 + 5):
        for world in (1, 2, 3, 4, 8):
            shards = [ShardedMatrix.row_shard(rows, r, world) for r in range(world)]
            assert shards[0].start == 0 and sum(s.count for s in shards) == rows
            for a, b in zip(shards, shards[1:]):
                assert a.start + a.count == b.start and (b.start % ROW_BLOCK == 0 or b.count == 0)
    assert [ShardedMatrix.row_shard(32768, r, 8).count for r in range(8)] == [4096] * 8      # BASELINE configs[4]


def _slice_partial(orc, a, start, is_max):
    """What trn_arg{max,min}_slice_f32_dev reports for one slice (see reduce.cu / parallel.py)."""
    from oracle import SCALAR
    from trueno_b200.parallel import NO_CANDIDATE
    ident = -np.inf if is_max else np.inf
    if start == 0:
        i = orc.argmax(a, backend=SCALAR) if is_max else orc.argmin(a, backend=SCALAR)
        return f32(a[i]), i
    ok = ~np.isnan(a) & (a != ident)
    if not ok.any():
        return f32(ident), NO_CANDIDATE
    v = a[ok].max() if is_max else a[ok].min()
    return f32(v), start + int(np.flatnonzero(a == v)[0])


def _worker(rank, world, port, cases, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import oracle
    from trueno_b200 import parallel as par
    r, _, w = par.init_distributed("gloo")
    assert (r, w) == (rank, world) and par.world_size() == world
    orc = oracle.get()
    out = []
    for a, b in cases:
        sh = par.shard_range(a.size, rank, world, align=4)
        la, lb = a[sh.start:sh.start + sh.count], b[sh.start:sh.start + sh.count]
        # sum-type partials -> all_reduce(SUM)
        dot = par.combine_sum(torch.tensor([float(orc.dot(la, lb)) if sh.count else 0.0], dtype=torch.float32))
        ssq = par.combine_sum(torch.tensor([float(orc.dot(la, la)) if sh.count else 0.0], dtype=torch.float32)).sqrt()
        mx = par.combine_extreme(torch.tensor([la.max() if sh.count else -np.inf], dtype=torch.float32), True)
        res = [float(dot), float(ssq), float(mx)]
        for is_max in (True, False):
            if sh.count:
                v, i = _slice_partial(orc, la, sh.start, is_max)
            else:
                v, i = f32(-np.inf if is_max else np.inf), par.NO_CANDIDATE
            gv, gi = par.combine_arg(torch.tensor([v], dtype=torch.float32), torch.tensor([i], dtype=torch.int64), is_max)
            res += [float(gv), int(gi)]
        # ONE softmax vector over the ranks: all_gather of (max, sum of exp) pairs in rank order
        with np.errstate(all="ignore"):
            if sh.count:
                m = f32(np.fmax.reduce(la))
                pair = [m, f32(np.sum(np.exp((la - m).astype(f32)), dtype=np.float64))]
            else:
                pair = [f32(-np.inf), f32(0)]
        pairs = par.gather_softmax_pairs(torch.tensor(pair, dtype=torch.float32))
        assert pairs.shape == (world, 2)
        res.append(pairs.numpy().astype(np.float64).tolist())
        out.append(res)
    if rank == 0:
        results.put(out)
    dist.barrier()
    dist.destroy_process_group()


def _cases():
    rng = np.random.default_rng(42)
    cases = []
    for n in (8, 9, 1000, 4097):
        a = rng.integers(-3, 4, n).astype(f32)          # heavy ties across slices
        cases.append((a, rng.standard_normal(n).astype(f32)))
    a = rng.standard_normal(64).astype(f32); a[0] = np.nan                       # NaN seed: index 0 wins
    cases.append((a, np.ones(64, f32)))
    a = rng.standard_normal(64).astype(f32); a[32] = np.nan; a[40] = 9; a[50] = 9  # NaN heads slice 1, max behind it
    cases.append((a, np.ones(64, f32)))
    a = np.full(64, -np.inf, f32); a[33:40] = np.nan                              # nothing beats the identity -> 0
    cases.append((a, np.ones(64, f32)))
    a = np.zeros(64, f32); a[5] = 7; a[37] = 7                                    # equal maxima in two slices
    cases.append((a, np.ones(64, f32)))
    return cases


@pytest.mark.parametrize("world", [2, 3])
def test_exchange_steps_over_gloo(world):
    import oracle
    from oracle import SCALAR
    from trueno_b200.parallel import shard_range
    orc = oracle.get()
    cases = _cases()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, cases, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get()
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for (a, b), res in zip(cases, got):
        dot, nrm, mx, vmax, imax, vmin, imin, pairs = res
        # the gathered pairs are the slices' pairs in rank order, and their fold is the whole vector's (max, sum of exp)
        with np.errstate(all="ignore"):
            fold_m, fold_s = -np.inf, 0.0
            for r, (pm, ps) in enumerate(pairs):
                sh = shard_range(a.size, r, world, align=4)
                la = a[sh.start:sh.start + sh.count]
                if sh.count and not np.isnan(la).all():
                    assert pm == np.fmax.reduce(la) or (np.isnan(pm) and np.isnan(np.fmax.reduce(la)))
                mn = max(fold_m, pm) if not np.isnan(pm) else fold_m
                ref = 0.0 if mn == -np.inf else mn
                fold_s = fold_s * np.exp(fold_m - ref) + ps * np.exp(pm - ref)
                fold_m = mn
            if np.isfinite(a).all():
                tm = a.max()
                ts = np.sum(np.exp((a - tm).astype(np.float64)))
                assert fold_m == tm and abs(fold_s - ts) <= 1e-6 * ts
        tdot, adot = orc.f64_dot(a, b)
        if np.isfinite(tdot):
            assert abs(dot - tdot) <= 1e-5 * adot + 1e-30
        with np.errstate(invalid="ignore"):
            tn = np.sqrt(np.sum(a.astype(np.float64) ** 2))
        if np.isfinite(tn):
            assert abs(nrm - tn) <= 1e-5 * tn + 1e-30
        assert imax == orc.argmax(a, backend=SCALAR), (a, imax)
        assert imin == orc.argmin(a, backend=SCALAR), (a, imin)
        wmax, wmin = orc.max(a, backend=SCALAR), orc.min(a, backend=SCALAR)
        assert (np.isnan(vmax) and np.isnan(wmax)) or vmax == wmax
        assert (np.isnan(vmin) and np.isnan(wmin)) or vmin == wmin


def _ragged_worker(rank, world, port, counts):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from trueno_b200 import parallel as par
    par.init_distributed("gloo")
    local = torch.arange(counts[rank], dtype=torch.float32) + 100 * rank
    out = par._gather_ragged(local, counts)
    want = torch.cat([torch.arange(c, dtype=torch.float32) + 100 * r for r, c in enumerate(counts)])
    assert torch.equal(out, want), (out, want)
    # the host-side contract of the sharded containers: an EMPTY sharded vector is the reference's error on every rank
    # alike (src/vector.rs:750-752, :1517-1519), before anything is launched or exchanged
    import trueno_b200 as trn
    empty = par.ShardedVector.__new__(par.ShardedVector)
    empty.local, empty.shard, empty.comm = torch.empty(0), par.shard_range(0, rank, world, 4), None
    for op, err in (("argmax", trn.TruenoError.InvalidInput("Empty vector")), ("min", trn.TruenoError.InvalidInput("Empty vector")),
                    ("softmax", trn.TruenoError.EmptyVector)):
        try:
            getattr(empty, op)()
            raise AssertionError(op)
        except trn.TruenoError as e:
            assert e == err
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("counts", [[5, 0, 2], [0, 0, 3], [4, 4]])
def test_ragged_gather_and_empty_vector_contract_over_gloo(counts):
    """ShardedMatrix.gather / gather_vector: slices of different (and zero) lengths, one padded all_gather."""
    world = len(counts)
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_ragged_worker, args=(r, world, port, counts)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0


def _peer_worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from trueno_b200 import parallel as par
    comm = par.PeerComm.try_create()          # no GPU here: no rank can export a mailbox
    raised = False
    try:
        par.PeerComm()
    except Exception:
        raised = True
    dist.barrier()                             # both ranks got here: nobody was left inside a collective
    if rank == 0:
        results.put((comm is None, raised))
    dist.destroy_process_group()


def test_peer_comm_fails_on_every_rank_together_without_a_gpu():
    """PeerComm's constructor keeps every rank inside the same collectives whatever fails locally, so a box without a P2P
    path (here: without a GPU) yields the same error on all ranks — `try_create()` returns None and the sharded reductions
    fall back to the NCCL exchange — instead of one rank raising while the others wait in an all_gather."""
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the fused path is exercised by the -m gpu tests")
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_peer_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get() == (True, True)
