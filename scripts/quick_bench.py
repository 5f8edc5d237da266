"""Development micro-benchmark: times each hot-path kernel device-resident with CUDA events.
Not the judged bench (that is bench.py) — this is the per-kernel exploration tool.
usage: python scripts/quick_bench.py [reduce] [map] [softmax] [gemm] [batched] [matvec]"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trueno_b200 as trn  # noqa: E402

PEAKS = {}
try:
    PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
except Exception:
    pass
HBM = PEAKS.get("hbm_gbs", 6650.0)


def timeit(fn, iters=20, warmup=3):
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
    which = set(sys.argv[1:]) or {"reduce", "map", "softmax", "gemm", "batched", "matvec"}   # + rowblock, hostbatched on request
    torch.cuda.set_device(0)
    trn.check(trn.lib.trn_cuda_init(0))
    # a real (non-default) stream: torch events and our launches must share it
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    assert st != 0
    L = trn.lib
    print(trn.device_info(), "HBM peak (measured)", HBM)

    if "reduce" in which:
        n = 1 << 30
        a = torch.rand(n, device="cuda") * 2 - 1
        b = torch.rand(n, device="cuda") * 2 - 1
        out = torch.zeros(4, device="cuda")
        oidx = torch.zeros(2, device="cuda", dtype=torch.int64)
        for name, fn, nbytes in [
            ("sum", lambda: trn.check(L.trn_sum_f32_dev(a.data_ptr(), n, out.data_ptr(), st)), 4 * n),
            ("dot", lambda: trn.check(L.trn_dot_f32_dev(a.data_ptr(), n, b.data_ptr(), n, out.data_ptr(), st)), 8 * n),
            ("norm_l2", lambda: trn.check(L.trn_norm_l2_f32_dev(a.data_ptr(), n, out.data_ptr(), st)), 4 * n),
            ("argmax", lambda: trn.check(L.trn_argmax_f32_dev(a.data_ptr(), n, oidx.data_ptr(), out.data_ptr(), st)), 4 * n),
            ("max", lambda: trn.check(L.trn_max_f32_dev(a.data_ptr(), n, out.data_ptr(), st)), 4 * n),
            ("torch.sum (ref)", lambda: torch.sum(a), 4 * n),
        ]:
            med, best = timeit(fn)
            print(f"reduce {name:16s} n=2^30 median {med:.3f} ms best {best:.3f} ms  {nbytes / med / 1e6:.0f} GB/s "
                  f"({nbytes / med / 1e6 / HBM:.2%} of measured HBM)")
        del a, b

    if "map" in which:
        n = 4096 * 32000
        a = torch.randn(n, device="cuda")
        b = torch.randn(n, device="cuda")
        o = torch.empty(n, device="cuda")
        for name, fn, nbytes in [
            ("add", lambda: trn.check(L.trn_add_f32_dev(a.data_ptr(), n, b.data_ptr(), n, o.data_ptr(), st)), 12 * n),
            ("mul", lambda: trn.check(L.trn_mul_f32_dev(a.data_ptr(), n, b.data_ptr(), n, o.data_ptr(), st)), 12 * n),
            ("sigmoid", lambda: trn.check(L.trn_sigmoid_f32_dev(a.data_ptr(), n, o.data_ptr(), st)), 8 * n),
            ("gelu", lambda: trn.check(L.trn_gelu_f32_dev(a.data_ptr(), n, o.data_ptr(), st)), 8 * n),
            ("torch.add (ref)", lambda: torch.add(a, b, out=o), 12 * n),
        ]:
            med, best = timeit(fn)
            print(f"map {name:16s} n=131M median {med:.3f} ms best {best:.3f} ms  {nbytes / med / 1e6:.0f} GB/s "
                  f"({nbytes / med / 1e6 / HBM:.2%})")
        del a, b, o

    if "softmax" in which:
        rows, cols = 4096, 32000
        a = torch.randn(rows, cols, device="cuda") * 4
        o = torch.empty_like(a)
        nbytes = 8 * rows * cols
        for name, fn in [
            ("softmax", lambda: trn.check(L.trn_softmax_rows_f32_dev(a.data_ptr(), o.data_ptr(), rows, cols, st))),
            ("log_softmax", lambda: trn.check(L.trn_log_softmax_rows_f32_dev(a.data_ptr(), o.data_ptr(), rows, cols, st))),
            ("torch.softmax (ref)", lambda: torch.softmax(a, dim=1, out=o)),
        ]:
            med, best = timeit(fn)
            print(f"softmax {name:20s} 4096x32000 median {med:.3f} ms best {best:.3f} ms  {nbytes / med / 1e6:.0f} GB/s "
                  f"({nbytes / med / 1e6 / HBM:.2%})")
        del a, o

    if "gemm" in which:
        for size in (2048, 4096, 8192):
            m = k = n = size
            a = torch.rand(m, k, device="cuda")
            b = torch.rand(k, n, device="cuda")
            c = torch.empty(m, n, device="cuda")
            flop = 2.0 * m * n * k
            for name, eng in [("simt", 1), ("tc 3xTF32", 2), ("tc 1xTF32", 3)]:
                if eng == 1 and size > 4096:
                    continue
                trn.set_gemm_engine(eng)
                fn = lambda: trn.check(L.trn_matmul_f32_dev(a.data_ptr(), m, k, b.data_ptr(), k, n, c.data_ptr(), st))
                med, best = timeit(fn, iters=10)
                print(f"gemm {size}^3 {name:10s} median {med:.3f} ms best {best:.3f} ms  {flop / med / 1e9:.1f} TFLOP/s")
                if eng == 2:
                    ref = (a.double() @ b.double())
                    err = ((c.double() - ref).abs() / (a.double().abs() @ b.double().abs())).max().item()
                    print(f"      3xTF32 max err / sum|a||b| = {err:.3e}")
            trn.set_gemm_engine(0)
            torch.backends.cuda.matmul.allow_tf32 = True
            med, best = timeit(lambda: torch.matmul(a, b, out=c), iters=10)
            print(f"gemm {size}^3 cuBLAS TF32  median {med:.3f} ms best {best:.3f} ms  {flop / med / 1e9:.1f} TFLOP/s  (peak probe)")
            torch.backends.cuda.matmul.allow_tf32 = False
            med, best = timeit(lambda: torch.matmul(a, b, out=c), iters=5)
            print(f"gemm {size}^3 cuBLAS FP32  median {med:.3f} ms best {best:.3f} ms  {flop / med / 1e9:.1f} TFLOP/s")
            del a, b, c

    if "batched" in which:
        B, H, m, k, n = 8, 32, 2048, 128, 2048
        a = torch.rand(B * H * m * k, device="cuda")
        b = torch.rand(B * H * k * n, device="cuda")
        c = torch.empty(B * H * m * n, device="cuda")
        flop = 2.0 * B * H * m * n * k
        nbytes = 4.0 * (a.numel() + b.numel() + c.numel())
        for name, eng in [("tc 3xTF32", 2)]:
            trn.set_gemm_engine(eng)
            fn = lambda: trn.check(L.trn_batched_matmul_4d_f32_dev(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(),
                                                                   c.data_ptr(), B, H, m, k, n, st))
            med, best = timeit(fn, iters=10)
            print(f"batched4d {name} median {med:.3f} ms best {best:.3f} ms  {flop / med / 1e9:.1f} TFLOP/s  "
                  f"{nbytes / med / 1e6:.0f} GB/s algorithmic")
        trn.set_gemm_engine(0)
        a3, b3 = a.view(B * H, m, k), b.view(B * H, k, n)
        med, best = timeit(lambda: torch.bmm(a3, b3, out=c.view(B * H, m, n)), iters=5)
        print(f"batched4d cuBLAS FP32 median {med:.3f} ms  {flop / med / 1e9:.1f} TFLOP/s")
        del a, b, c

    if "attention" in which:
        # BASELINE config 3's head shape, now fused: out = softmax(scale * Q K^T) V per head
        H, seq, d = int(os.environ.get("ATT_HEADS", 256)), int(os.environ.get("ATT_SEQ", 2048)), int(os.environ.get("ATT_D", 128))
        q = torch.randn(H * seq * d, device="cuda"); k = torch.randn(H * seq * d, device="cuda"); v = torch.randn(H * seq * d, device="cuda")
        o = torch.empty(H * seq * d, device="cuda")
        scale = 1.0 / d ** 0.5
        for causal in (0, 1):
            flop = 4.0 * H * seq * seq * d * (0.5 if causal else 1.0)
            fn = lambda: trn.check(L.trn_attention_f32_dev(q.data_ptr(), q.numel(), k.data_ptr(), k.numel(), v.data_ptr(), v.numel(),
                                                           o.data_ptr(), H, seq, d, scale, causal, st))
            med, best = timeit(fn, iters=10)
            print(f"attention H={H} seq={seq} d={d} causal={causal} fused 3xTF32 median {med:.3f} ms best {best:.3f} ms  {flop / med / 1e9:.1f} TFLOP/s")
            q4, k4, v4 = (t.view(1, H, seq, d) for t in (q, k, v))
            torch.backends.cuda.matmul.allow_tf32 = False
            med, best = timeit(lambda: torch.nn.functional.scaled_dot_product_attention(q4, k4, v4, is_causal=bool(causal)), iters=5)
            print(f"   torch SDPA fp32 median {med:.3f} ms  {flop / med / 1e9:.1f} TFLOP/s")
        # unfused composition on this library's own kernels (needs the seq x seq scores in HBM)
        if H * seq * seq * 4 <= (8 << 30):
            s_buf = torch.empty(H * seq * seq, device="cuda")
            p_buf = torch.empty(H * seq * seq, device="cuda")
            kt = k.view(H, seq, d).transpose(1, 2).contiguous().view(-1)
            def unfused():
                trn.check(L.trn_batched_matmul_f32_dev(q.data_ptr(), q.numel(), kt.data_ptr(), kt.numel(), s_buf.data_ptr(), H, seq, d, seq, st))
                trn.check(L.trn_scale_f32_dev(s_buf.data_ptr(), s_buf.numel(), scale, s_buf.data_ptr(), st))
                trn.check(L.trn_softmax_rows_f32_dev(s_buf.data_ptr(), p_buf.data_ptr(), H * seq, seq, st))
                trn.check(L.trn_batched_matmul_f32_dev(p_buf.data_ptr(), p_buf.numel(), v.data_ptr(), v.numel(), o.data_ptr(), H, seq, seq, d, st))
            med, best = timeit(unfused, iters=5)
            print(f"   unfused (bmm -> scale -> softmax -> bmm, K^T given) median {med:.3f} ms  {4.0 * H * seq * seq * d / med / 1e9:.1f} TFLOP/s")

    if "eigen" in which:
        import time
        import numpy as np
        for n in (64, 256, 1024, 2048, 4096):
            rng = np.random.default_rng(n)
            x = rng.standard_normal((n, n)).astype(np.float32)
            m = torch.from_numpy(((x + x.T) / 2).astype(np.float32)).cuda()
            vals = torch.empty(n, device="cuda"); vecs = torch.empty(n * n, device="cuda")
            fn = lambda: trn.check(L.trn_symmetric_eigen_f32_dev(m.data_ptr(), n, n, vals.data_ptr(), vecs.data_ptr(), st))
            fn(); torch.cuda.synchronize()
            l0 = trn.launch_count() if hasattr(trn, "launch_count") else 0
            t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
            l1 = trn.launch_count() if hasattr(trn, "launch_count") else 0
            t1 = time.perf_counter(); w = torch.linalg.eigvalsh(m.double()); torch.cuda.synchronize(); dt_t = time.perf_counter() - t1
            t1 = time.perf_counter(); w32, v32 = torch.linalg.eigh(m); torch.cuda.synchronize(); dt_t32 = time.perf_counter() - t1
            err = float((vals.double().flip(0) - w).abs().max() / m.double().norm())
            rounds = (l1 - l0) / max(1, n - 1 + n % 2)
            print(f"symmetric_eigen n={n}: {dt * 1e3:.1f} ms  (~{rounds:.1f} sweeps, {l1 - l0} launches)  max eigenvalue error {err:.2e} * ||A||_F   "
                  f"torch eigh f32 (cuSOLVER) {dt_t32 * 1e3:.1f} ms")

    if "rowblock" in which:
        # BASELINE config 5b: one GPU's share of the 32768^3 product at 8 GPUs (A-block 4096 x 32768, full B)
        n, mb = 32768, 4096
        a = torch.rand(mb, n, device="cuda")
        b = torch.rand(n, n, device="cuda")
        c = torch.empty(mb, n, device="cuda")
        flop = 2.0 * mb * n * n
        fn = lambda: trn.check(L.trn_matmul_f32_dev(a.data_ptr(), mb, n, b.data_ptr(), n, n, c.data_ptr(), st))
        med, best = timeit(fn, iters=5)
        print(f"rowblock 4096x32768x32768 tc 3xTF32 median {med:.3f} ms best {best:.3f} ms  {flop / med / 1e9:.1f} TFLOP/s")
        del a, b, c

    if "hostbatched" in which:
        # BASELINE config 3 through the host-slice API (pinned): 512 MiB up, 4 GiB down, pipelined by head groups
        B, H, m, k, n = 8, 32, 2048, 128, 2048
        ha, hb, hc = trn.pinned_empty(B * H * m * k), trn.pinned_empty(B * H * k * n), trn.pinned_empty(B * H * m * n)
        ha[:] = 0.5
        hb[:] = 0.25
        import time
        fn = lambda: trn.check(L.trn_batched_matmul_4d_f32(ha.ctypes.data, ha.size, hb.ctypes.data, hb.size, hc.ctypes.data, B, H, m, k, n))
        fn()
        t0 = time.perf_counter()
        for _ in range(3):
            fn()
        dt = (time.perf_counter() - t0) / 3
        print(f"host batched4d (pinned, pipelined) {dt * 1e3:.1f} ms  {2.0 * B * H * m * n * k / dt / 1e12:.1f} TFLOP/s e2e  D2H {4.0 * hc.size / dt / 1e9:.1f} GB/s  check {float(hc[12345]):.3f}")
        del ha, hb, hc

    if "matvec" in which:
        rows = cols = 16384
        a = torch.randn(rows, cols, device="cuda")
        v = torch.randn(cols, device="cuda")
        y = torch.empty(rows, device="cuda")
        med, best = timeit(lambda: trn.check(L.trn_matvec_f32_dev(a.data_ptr(), rows, cols, v.data_ptr(), cols, y.data_ptr(), st)))
        print(f"matvec 16384^2 median {med:.3f} ms  {4.0 * rows * cols / med / 1e6:.0f} GB/s ({4.0 * rows * cols / med / 1e6 / HBM:.2%})")
    print("launches:", trn.launch_count())


if __name__ == "__main__":
    main()
