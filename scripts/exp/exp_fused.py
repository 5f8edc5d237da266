"""Fused-split GEMM vs the pre-pass GEMM: config 3 and a K sweep.  Run once per TRN_GEMM_FUSED mode (read once per process)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import trueno_b200 as trn  # noqa: E402


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in evs:
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    torch.cuda.set_device(0)
    trn.check(trn.lib.trn_cuda_init(0))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    L = trn.lib
    mode = os.environ.get("TRN_GEMM_FUSED", "auto")
    which = sys.argv[1:] or ["cfg3", "sweep"]
    if "cfg3" in which:
        B, H, S, D = 8, 32, 2048, 128
        q = torch.rand(B * H * S * D, device="cuda")
        kt = torch.rand(B * H * D * S, device="cuda")
        c = torch.empty(B * H * S * S, device="cuda")
        fn = lambda: trn.check(L.trn_batched_matmul_4d_f32_dev(q.data_ptr(), q.numel(), kt.data_ptr(), kt.numel(), c.data_ptr(), B, H, S, D, S, st))
        med, best = timeit(fn)
        flop = 2.0 * B * H * S * S * D
        print(f"[fused={mode}] config3 batched_matmul_4d 8x32x2048x128x2048: median {med:.3f} ms best {best:.3f} ms = {flop / med / 1e9:.1f} TF/s")
        # quick parity: one head against f64
        h = 77
        truth = q.view(B * H, S, D)[h].double() @ kt.view(B * H, D, S)[h].double()
        err = (c.view(B * H, S, S)[h].double() - truth).abs().max().item() / D
        print(f"   head {h} max |err| / K = {err:.3e}")
        del q, kt, c
    if "sweep" in which:
        for (m, k, n) in [(8192, 128, 8192), (8192, 256, 8192), (8192, 512, 8192), (8192, 1024, 8192), (8192, 2048, 8192),
                          (8192, 8192, 8192), (4096, 4096, 4096), (2048, 2048, 2048), (4096, 32768, 4096)]:
            a = torch.rand(m * k, device="cuda")
            b = torch.rand(k * n, device="cuda")
            c = torch.empty(m * n, device="cuda")
            fn = lambda: trn.check(L.trn_matmul_f32_dev(a.data_ptr(), m, k, b.data_ptr(), k, n, c.data_ptr(), st))
            med, best = timeit(fn, iters=6, warmup=2)
            print(f"[fused={mode}] matmul {m}x{k}x{n}: median {med:.3f} ms best {best:.3f} = {2.0 * m * k * n / med / 1e9:.1f} TF/s")
            del a, b, c


main()
