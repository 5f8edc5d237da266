"""Multi-GPU parity worker: run under torchrun (one rank per GPU, NCCL).  Every rank regenerates the
SAME global inputs from a seed, computes its shard through the C-ABI `_dev` kernels + the exchange
steps of trueno_b200/parallel.py, and checks the result against the CPU oracle on the global input
(slices + all_reduce / all_gather; head ranges; row blocks).  Exit code 0 == pass."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
f32 = np.float32


def main() -> int:
    import oracle
    import trueno_b200 as trn
    from oracle import SCALAR
    from trueno_b200 import parallel as par

    rank, local_rank, world = par.init_distributed("nccl")
    torch.cuda.set_device(local_rank)
    trn.check(trn.lib.trn_cuda_init(local_rank))
    dev = torch.device("cuda", local_rank)
    orc = oracle.get()
    st = par.current_stream_handle()
    rng = np.random.default_rng(2026)

    # ---- config 4 (scaled to what the oracle finishes in seconds): sliced reductions, exchanged (a) by NCCL
    #      behind the slice kernel and (b) inside the slice kernel over NVLink peer memory (csrc/peer.cu)
    comm = par.PeerComm()
    for n, use_comm in [(x, c) for x in (1 << 22, (1 << 22) + 37, 1001) for c in (None, comm)]:
        a = rng.uniform(-1, 1, n).astype(f32)
        b = rng.uniform(-1, 1, n).astype(f32)
        # planted maximum in the LAST rank's slice and an equal duplicate later in the same slice,
        # plus an equal value in slice 0 at a lower index: the lowest global index must win
        sh_last = par.shard_range(n, world - 1, world, align=4)
        big = f32(a.max() + 1)
        a[sh_last.start + 3] = big
        a[min(sh_last.start + 11, n - 1)] = big
        a[5] = big
        a[7] = f32(a.min() - 1)
        sh = par.shard_range(n, rank, world, align=4)
        va = par.ShardedVector(torch.from_numpy(a[sh.start:sh.start + sh.count]).to(dev), sh, use_comm)
        vb = par.ShardedVector(torch.from_numpy(b[sh.start:sh.start + sh.count]).to(dev), sh, use_comm)
        tdot, adot = orc.f64_dot(a, b)
        tsum, asum = orc.f64_sum(a)
        assert abs(float(va.dot(vb)) - tdot) <= 1e-5 * adot
        assert abs(float(va.sum()) - tsum) <= 1e-5 * asum
        tn = float(np.sqrt(np.sum(a.astype(np.float64) ** 2)))
        assert abs(float(va.norm_l2()) - tn) <= 1e-5 * tn
        assert int(va.argmax()) == orc.argmax(a, backend=SCALAR) == 5
        assert int(va.argmin()) == orc.argmin(a, backend=SCALAR) == 7
        assert float(va.max()) == float(big)
        if use_comm is not None:
            # every rank folds the slice results in rank order: the fused answers are bit-identical everywhere
            mine = torch.stack([va.dot(vb).clone(), va.sum().clone(), va.norm_l2().clone()]).reshape(-1)
            everyone = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(everyone, mine)
            assert all(torch.equal(everyone[0], e) for e in everyone)
            # NaN seed in slice 0 wins across ranks; repeated calls stay in step (sequence numbers)
            a2 = a.copy(); a2[0] = np.nan
            vn = par.ShardedVector(torch.from_numpy(a2[sh.start:sh.start + sh.count]).to(dev), sh, use_comm)
            for _ in range(5):
                assert int(vn.argmax()) == 0 and int(vn.argmin()) == 0

    # ---- fewer aligned blocks than ranks: some slices are EMPTY and still take part in every exchange (no rank left waiting)
    for n, use_comm in [(x, c) for x in (3, 5, 9) for c in (None, comm)]:
        a = np.arange(1, n + 1, dtype=f32)
        a[n - 1] = -7
        sh = par.shard_range(n, rank, world, align=4)
        va = par.ShardedVector(torch.from_numpy(a[sh.start:sh.start + sh.count]).to(dev), sh, use_comm)
        assert float(va.sum()) == float(a.sum()) and int(va.argmin()) == n - 1 and int(va.argmax()) == n - 2
        assert float(va.min()) == -7.0 and float(va.max()) == float(n - 1)
        got = va.softmax()
        torch.cuda.synchronize()
        e = np.exp((a - a.max()).astype(np.float64))
        assert np.allclose(got.cpu().numpy(), (e / e.sum())[sh.start:sh.start + sh.count], atol=1e-6)
    empty = par.ShardedVector(torch.empty(0, device=dev), par.shard_range(0, rank, world, 4), comm)
    for op in ("argmax", "min"):
        try:
            getattr(empty, op)()
            raise AssertionError("empty vector must raise")
        except trn.TruenoError as err:
            assert err == trn.TruenoError.InvalidInput("Empty vector")
    assert float(empty.sum()) == 0.0            # src/vector.rs:635: the sum of an empty vector is 0.0

    # ---- a chain of fused reductions captured in a CUDA graph and replayed (device-side call numbers)
    n = 1 << 20
    a = rng.uniform(-1, 1, n).astype(f32)
    sh = par.shard_range(n, rank, world, align=4)
    stream = torch.cuda.Stream()
    with torch.cuda.stream(stream):
        va = par.ShardedVector(torch.from_numpy(a[sh.start:sh.start + sh.count]).to(dev), sh, comm)
        loop = par.CapturedLoop(lambda: (va.sum(), va.norm_l2(), va.argmax()), 6)
        for _ in range(3):
            loop.replay()
        stream.synchronize()
        tsum, asum = orc.f64_sum(a)
        assert float(va._f32) == float(a.max()) and int(va._i64) == orc.argmax(a, backend=SCALAR)   # last call of the chain: argmax
        assert abs(float(va.sum()) - tsum) <= 1e-5 * asum
    trn.check(trn.lib.trn_comm_status(comm.handle))

    # ---- entry points called from a worker thread of a process bound to device `local_rank` (per-thread current device)
    import threading
    box = {}

    def from_thread():
        v = trn.Vector.from_slice(np.arange(1000, dtype=f32))
        box["sum"] = float(v.sum())
        box["mm"] = trn.Matrix.from_vec(2, 2, [1, 2, 3, 4]).matmul(trn.Matrix.from_vec(2, 2, [5, 6, 7, 8])).to_numpy().tolist()
    t = threading.Thread(target=from_thread)
    t.start()
    t.join()
    assert box == {"sum": 499500.0, "mm": [[19.0, 22.0], [43.0, 50.0]]}, box

    # ---- ONE softmax / log_softmax vector sharded over the ranks: slice stats -> all_gather of pairs -> fold + write
    for n in (1 << 22, (1 << 22) + 37, 1001):
        a = (rng.standard_normal(n) * 4).astype(f32)
        a[n - 3] = 11.0          # the maximum lives in the last rank's slice
        sh = par.shard_range(n, rank, world, align=4)
        va = par.ShardedVector(torch.from_numpy(a[sh.start:sh.start + sh.count]).to(dev), sh)
        arg = (a - a.max()).astype(f32).astype(np.float64)
        e64 = np.exp(arg)
        for log in (False, True):
            got = va.softmax(log=log)
            torch.cuda.synchronize()
            g64 = got.cpu().numpy().astype(np.float64)
            want = (arg - np.log(e64.sum()) if log else e64 / e64.sum())[sh.start:sh.start + sh.count]
            ulp = np.spacing(np.abs(want).astype(f32)).astype(np.float64)
            if log:
                assert np.all(np.abs(g64 - want) <= 4 * ulp + 2.0 ** -20)
            else:
                assert np.all(np.abs(g64 - want) <= np.minimum(1e-6, 8 * ulp + 1e-45))
            # slices of every rank add up to one; the same bits normalise every slice
            tot = torch.tensor([g64.sum() if not log else 0.0], dtype=torch.float64, device=dev)
            dist.all_reduce(tot)
            if not log:
                assert abs(float(tot) - 1.0) < 1e-5

    # ---- config 3 (scaled): batch x head ranges, no collective; every rank checks its own heads ---------
    B, H, m, k, n = 2, 8, 256, 128, 384
    A = rng.uniform(-1, 1, (B * H, m, k)).astype(f32)
    Bm = rng.uniform(-1, 1, (B * H, k, n)).astype(f32)
    hs = par.shard_range(B * H, rank, world)
    dA = torch.from_numpy(A[hs.start:hs.start + hs.count]).to(dev).contiguous()
    dB = torch.from_numpy(Bm[hs.start:hs.start + hs.count]).to(dev).contiguous()
    dC = torch.empty(hs.count, m, n, device=dev)
    trn.check(trn.lib.trn_batched_matmul_f32_dev(dA.data_ptr(), dA.numel(), dB.data_ptr(), dB.numel(), dC.data_ptr(),
                                                 hs.count, m, k, n, st))
    torch.cuda.synchronize()
    got = dC.cpu().numpy()
    for i in range(hs.count):
        h = hs.start + i
        truth = A[h].astype(np.float64) @ Bm[h].astype(np.float64)
        scale = np.abs(A[h]).astype(np.float64) @ np.abs(Bm[h]).astype(np.float64)
        assert np.all(np.abs(got[i] - truth) <= 1e-5 * scale)

    # ---- config 5 (scaled): row-sharded softmax / log_softmax / gelu, and row-block matmul with B replicated
    rows, cols = 64, 32000
    X = (rng.standard_normal((rows, cols)) * 4).astype(f32)
    rs = par.shard_range(rows, rank, world)
    dX = torch.from_numpy(X[rs.start:rs.start + rs.count]).to(dev).contiguous()
    dY = torch.empty_like(dX)
    trn.check(trn.lib.trn_softmax_rows_f32_dev(dX.data_ptr(), dY.data_ptr(), rs.count, cols, st))
    torch.cuda.synchronize()
    want = orc.softmax_rows(X[rs.start:rs.start + rs.count], rs.count, cols, backend=SCALAR)
    assert np.all(np.abs(dY.cpu().numpy() - want) <= 1e-6 + 2e-4 * want)
    trn.check(trn.lib.trn_gelu_f32_dev(dX.data_ptr(), dX.numel(), dY.data_ptr(), st))
    torch.cuda.synchronize()
    gw = orc.gelu(X[rs.start:rs.start + rs.count].reshape(-1), backend=SCALAR)
    assert np.max(np.abs(dY.cpu().numpy().reshape(-1) - gw)) <= 2e-6

    M, K, N = 1024, 512, 768
    A2 = rng.uniform(-1, 1, (M, K)).astype(f32)
    B2 = rng.uniform(-1, 1, (K, N)).astype(f32)
    ms = par.shard_range(M, rank, world, align=128)
    dA2 = torch.from_numpy(A2[ms.start:ms.start + ms.count]).to(dev).contiguous()
    dB2 = torch.from_numpy(B2).to(dev)
    dC2 = torch.empty(ms.count, N, device=dev)
    trn.check(trn.lib.trn_matmul_f32_dev(dA2.data_ptr(), ms.count, K, dB2.data_ptr(), K, N, dC2.data_ptr(), st))
    # gather the row blocks (optional step of §8e) and compare the WHOLE product on every rank
    blocks = [torch.empty(par.shard_range(M, r, world, align=128).count, N, device=dev) for r in range(world)]
    dist.all_gather(blocks, dC2)
    full = torch.cat(blocks).cpu().numpy()
    truth = A2.astype(np.float64) @ B2.astype(np.float64)
    scale = np.abs(A2).astype(np.float64) @ np.abs(B2).astype(np.float64)
    assert np.all(np.abs(full - truth) <= 1e-5 * scale)

    # ---- ShardedMatrix: row blocks of whole 256-row tiles, B replicated by ONE broadcast and pre-split once; matvec by rows
    M, K, N = 1024 + 300, 640, 512
    A3 = rng.uniform(-1, 1, (M, K)).astype(f32)
    B3 = rng.uniform(-1, 1, (K, N)).astype(f32)
    v3 = rng.uniform(-1, 1, K).astype(f32)
    rs3 = par.ShardedMatrix.row_shard(M)
    assert rs3.start % par.ROW_BLOCK == 0 or rs3.count == 0
    Am = par.ShardedMatrix(torch.from_numpy(A3[rs3.start:rs3.start + rs3.count]).to(dev).reshape(-1).contiguous(), rs3, K)
    Bsrc = torch.from_numpy(B3).to(dev).reshape(-1) if rank == 0 else None          # only rank 0 holds B
    Bm = par.ReplicatedMatrix.broadcast(Bsrc, K, N, src=0, device=dev)
    plain = Am.matmul(Bm).gather()
    prepared = Am.matmul(Bm.prepare()).gather()
    torch.cuda.synchronize()
    assert torch.equal(plain, prepared)                       # the prepared handle changes no bit
    truth = A3.astype(np.float64) @ B3.astype(np.float64)
    scale = np.abs(A3).astype(np.float64) @ np.abs(B3).astype(np.float64)
    assert np.all(np.abs(plain.cpu().numpy() - truth) <= 1e-5 * scale)
    if rs3.count:                                             # and none against the unsharded product of the same rows
        whole = torch.empty(M * N, device=dev)
        dA = torch.from_numpy(A3).to(dev)
        trn.check(trn.lib.trn_matmul_f32_dev(dA.data_ptr(), M, K, Bm.data.data_ptr(), K, N, whole.data_ptr(), st))
        torch.cuda.synchronize()
        assert torch.equal(whole.view(M, N)[rs3.start:rs3.start + rs3.count], plain[rs3.start:rs3.start + rs3.count])
    y = par.gather_vector(Am.matvec(torch.from_numpy(v3).to(dev)))
    ty = A3.astype(np.float64) @ v3.astype(np.float64)
    sy = np.abs(A3).astype(np.float64) @ np.abs(v3).astype(np.float64)
    assert np.all(np.abs(y.cpu().numpy() - ty) <= 1e-5 * sy)
    try:
        Am.matmul(par.ReplicatedMatrix(torch.zeros(3 * 4, device=dev), 3, 4))
        raise AssertionError("inner dimensions must be checked")
    except trn.TruenoError as err:
        assert err.variant == "InvalidInput" and f"inner dimensions {K} and 3 must match" in err.message
    Bm.close()

    dist.barrier()
    if rank == 0:
        print(f"dist_worker ok: world={world}, launches={trn.launch_count()}")
    dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
