"""One launch of every hot-path kernel at its BASELINE size between cudaProfilerStart/Stop, for
    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_all python scripts/profile_kernels.py
Warm-up launches run before the profiler range so the captured launch is steady-state (apart from ncu's own
cache control).  usage: python scripts/profile_kernels.py [gemm] [batched] [reduce] [map] [softmax] [rows] [rows2] [matvec] [attention] [conv]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trueno_b200 as trn  # noqa: E402


def main():
    which = set(sys.argv[1:]) or {"gemm", "batched", "reduce", "map", "softmax", "rows", "matvec", "attention", "conv"}
    torch.cuda.set_device(0)
    trn.check(trn.lib.trn_cuda_init(0))
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream
    L = trn.lib
    ops = []
    keep = []
    if "gemm" in which:
        n = 8192
        a, b, c = torch.rand(n, n, device="cuda"), torch.rand(n, n, device="cuda"), torch.empty(n, n, device="cuda")
        keep += [a, b, c]
        ops.append(lambda: trn.check(L.trn_matmul_f32_dev(a.data_ptr(), n, n, b.data_ptr(), n, n, c.data_ptr(), st)))
    if "batched" in which:
        B, H, m, k, nn = 8, 32, 2048, 128, 2048
        qa, qb = torch.rand(B * H * m * k, device="cuda"), torch.rand(B * H * k * nn, device="cuda")
        qc = torch.empty(B * H * m * nn, device="cuda")
        keep += [qa, qb, qc]
        ops.append(lambda: trn.check(L.trn_batched_matmul_4d_f32_dev(qa.data_ptr(), qa.numel(), qb.data_ptr(), qb.numel(),
                                                                     qc.data_ptr(), B, H, m, k, nn, st)))
    if "fused" in which:
        # the fused-split GEMM (no pre-pass) at a K where it is the default but the A panel is not stationary (K = 256)
        fm, fk, fn_ = 8192, 256, 8192
        fa, fb, fc = torch.rand(fm, fk, device="cuda"), torch.rand(fk, fn_, device="cuda"), torch.empty(fm, fn_, device="cuda")
        keep += [fa, fb, fc]
        ops.append(lambda: trn.check(L.trn_matmul_f32_dev(fa.data_ptr(), fm, fk, fb.data_ptr(), fk, fn_, fc.data_ptr(), st)))
    if "slice" in which:
        # one GPU's slice of config 4 at 8 GPUs: 2^27 elements
        ns = 1 << 27
        sx_ = torch.rand(ns, device="cuda") * 2 - 1
        so_ = torch.zeros(4, device="cuda")
        si_ = torch.zeros(2, dtype=torch.int64, device="cuda")
        keep += [sx_, so_, si_]
        ops += [lambda: trn.check(L.trn_sum_f32_dev(sx_.data_ptr(), ns, so_.data_ptr(), st)),
                lambda: trn.check(L.trn_argmax_f32_dev(sx_.data_ptr(), ns, si_.data_ptr(), so_.data_ptr(), st))]
    if "reduce" in which:
        n1 = 1 << 30
        x, y = torch.rand(n1, device="cuda") * 2 - 1, torch.rand(n1, device="cuda") * 2 - 1
        out = torch.zeros(4, device="cuda")
        oi = torch.zeros(2, dtype=torch.int64, device="cuda")
        keep += [x, y, out, oi]
        ops += [lambda: trn.check(L.trn_sum_f32_dev(x.data_ptr(), n1, out.data_ptr(), st)),
                lambda: trn.check(L.trn_dot_f32_dev(x.data_ptr(), n1, y.data_ptr(), n1, out.data_ptr(), st)),
                lambda: trn.check(L.trn_norm_l2_f32_dev(x.data_ptr(), n1, out.data_ptr(), st)),
                lambda: trn.check(L.trn_argmax_f32_dev(x.data_ptr(), n1, oi.data_ptr(), out.data_ptr(), st))]
    if "map" in which or "softmax" in which:
        rows, cols = 4096, 32000
        z = torch.randn(rows, cols, device="cuda") * 4
        z2 = torch.randn(rows, cols, device="cuda")
        o = torch.empty_like(z)
        keep += [z, z2, o]
        if "map" in which:
            ops += [lambda: trn.check(L.trn_add_f32_dev(z.data_ptr(), z.numel(), z2.data_ptr(), z2.numel(), o.data_ptr(), st)),
                    lambda: trn.check(L.trn_sigmoid_f32_dev(z.data_ptr(), z.numel(), o.data_ptr(), st)),
                    lambda: trn.check(L.trn_gelu_f32_dev(z.data_ptr(), z.numel(), o.data_ptr(), st))]
        if "softmax" in which:
            ops += [lambda: trn.check(L.trn_softmax_rows_f32_dev(z.data_ptr(), o.data_ptr(), rows, cols, st)),
                    lambda: trn.check(L.trn_log_softmax_rows_f32_dev(z.data_ptr(), o.data_ptr(), rows, cols, st))]
    if "rows" in which:
        # the other row-kernel families: warp-per-row (cols 1024), CTA-per-row registers (cols 8192), layer_norm, transpose
        for rr, cc in ((131072, 1024), (16384, 8192)):
            w = torch.randn(rr, cc, device="cuda")
            wo = torch.empty_like(w)
            g, bb = torch.randn(cc, device="cuda"), torch.randn(cc, device="cuda")
            keep += [w, wo, g, bb]
            ops += [lambda w=w, wo=wo, rr=rr, cc=cc: trn.check(L.trn_softmax_rows_f32_dev(w.data_ptr(), wo.data_ptr(), rr, cc, st)),
                    lambda w=w, wo=wo, g=g, bb=bb, rr=rr, cc=cc: trn.check(L.trn_layer_norm_rows_f32_dev(w.data_ptr(), g.data_ptr(), cc, bb.data_ptr(), cc, 1e-5, wo.data_ptr(), rr, cc, st))]
        t_in = torch.randn(16384, 8192, device="cuda")
        t_out = torch.empty(8192, 16384, device="cuda")
        keep += [t_in, t_out]
        ops.append(lambda: trn.check(L.trn_transpose_f32_dev(t_in.data_ptr(), 16384, 8192, t_out.data_ptr(), st)))
    if "rows2" in which:
        # long / window / split row kernels: LLM-vocabulary rows (128 256 aligned; 50 257 and 100 003 not 16-byte aligned),
        # window forms of the warp / CTA / ring kernels, one large Vector::softmax
        for rr, cc in ((4096, 32000), (1046, 128256), (2670, 50257), (1342, 100003), (5461, 24576), (134083, 1001), (16386, 8191), (4194, 32001), (1, 1 << 27)):
            w = torch.randn(rr, cc, device="cuda")
            wo = torch.empty_like(w)
            keep += [w, wo]
            ops += [lambda w=w, wo=wo, rr=rr, cc=cc: trn.check(L.trn_softmax_rows_f32_dev(w.data_ptr(), wo.data_ptr(), rr, cc, st)),
                    lambda w=w, wo=wo, rr=rr, cc=cc: trn.check(L.trn_log_softmax_rows_f32_dev(w.data_ptr(), wo.data_ptr(), rr, cc, st))]
        w = torch.randn(8192, 16384, device="cuda")
        wo = torch.empty_like(w)
        g, bb = torch.randn(16384, device="cuda"), torch.randn(16384, device="cuda")
        keep += [w, wo, g, bb]
        ops.append(lambda: trn.check(L.trn_layer_norm_rows_f32_dev(w.data_ptr(), g.data_ptr(), 16384, bb.data_ptr(), 16384, 1e-5, wo.data_ptr(), 8192, 16384, st)))
        tab = torch.randn(50257, 768, device="cuda")
        tidx = torch.randint(0, 50257, (1 << 19,), device="cuda", dtype=torch.int64)
        tout = torch.empty(1 << 19, 768, device="cuda")
        keep += [tab, tidx, tout]
        ops.append(lambda: trn.check(L.trn_embedding_lookup_f32_dev(tab.data_ptr(), 50257, 768, tidx.data_ptr(), tidx.numel(), tout.data_ptr(), st)))
        ops.append(lambda: trn.check(L.trn_mish_f32_dev(w.data_ptr(), w.numel(), wo.data_ptr(), st)))
        ops.append(lambda: trn.check(L.trn_hardswish_f32_dev(w.data_ptr(), w.numel(), wo.data_ptr(), st)))
    if "matvec" in which:
        r = 16384
        ma, mv, my = torch.randn(r, r, device="cuda"), torch.randn(r, device="cuda"), torch.empty(r, device="cuda")
        keep += [ma, mv, my]
        ops.append(lambda: trn.check(L.trn_matvec_f32_dev(ma.data_ptr(), r, r, mv.data_ptr(), r, my.data_ptr(), st)))
    if "attention" in which:
        aH, aseq, ad = 256, 2048, 128
        aq, ak, av = (torch.randn(aH * aseq * ad, device="cuda") for _ in range(3))
        ao = torch.empty_like(aq)
        keep += [aq, ak, av, ao]
        for causal in (0, 1):
            ops.append(lambda causal=causal: trn.check(L.trn_attention_f32_dev(aq.data_ptr(), aq.numel(), ak.data_ptr(), ak.numel(),
                                                                               av.data_ptr(), av.numel(), ao.data_ptr(), aH, aseq, ad,
                                                                               1.0 / ad ** 0.5, causal, st)))
    if "conv" in which:
        img, ker = torch.randn(8192, 8192, device="cuda"), torch.randn(5, 5, device="cuda")
        co = torch.empty(8188 * 8188, device="cuda")
        keep += [img, ker, co]
        ops.append(lambda: trn.check(L.trn_convolve2d_f32_dev(img.data_ptr(), 8192, 8192, ker.data_ptr(), 5, 5, co.data_ptr(), st)))
    for _ in range(3):
        for f in ops:
            f()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    for f in ops:
        f()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
    print("profiled", len(ops), "ops;", trn.launch_count(), "launches")


if __name__ == "__main__":
    main()
