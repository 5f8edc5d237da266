"""Row-kernel sweep: softmax / log_softmax / layer_norm / transpose / matvec over a range of shapes (GB/s, device-resident)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trueno_b200 as trn

def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters

torch.cuda.set_device(0); trn.check(trn.lib.trn_cuda_init(0))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
L = trn.lib
for rows, cols in [(1 << 20, 128), (1 << 19, 256), (1 << 18, 512), (1 << 17, 1024), (1 << 16, 2048), (1 << 15, 4096), (1 << 14, 8192), (8192, 16384), (4096, 32000), (2048, 65536), (1024, 131072)]:
    x = torch.randn(rows, cols, device="cuda"); y = torch.empty_like(x)
    g = torch.randn(cols, device="cuda"); b = torch.randn(cols, device="cuda")
    nb = 8.0 * rows * cols
    t1 = timeit(lambda: trn.check(L.trn_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
    t2 = timeit(lambda: trn.check(L.trn_log_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
    t3 = timeit(lambda: trn.check(L.trn_layer_norm_rows_f32_dev(x.data_ptr(), g.data_ptr(), cols, b.data_ptr(), cols, 1e-5, y.data_ptr(), rows, cols, st)))
    t4 = timeit(lambda: trn.check(L.trn_transpose_f32_dev(x.data_ptr(), rows, cols, y.data_ptr(), st)))
    t5 = timeit(lambda: torch.softmax(x, dim=1, out=y))
    print(f"{rows:8d} x {cols:6d}: softmax {nb/t1/1e6:6.0f}  log_softmax {nb/t2/1e6:6.0f}  layer_norm {nb/t3/1e6:6.0f}  transpose {nb/t4/1e6:6.0f}  torch.softmax {nb/t5/1e6:6.0f} GB/s")
    del x, y
for rows, cols in [(16384, 16384), (65536, 4096), (4096, 65536), (1 << 20, 256), (256, 1 << 20)]:
    a = torch.randn(rows, cols, device="cuda"); v = torch.randn(cols, device="cuda"); o = torch.empty(rows, device="cuda")
    w = torch.randn(rows, device="cuda"); o2 = torch.empty(cols, device="cuda")
    t = timeit(lambda: trn.check(L.trn_matvec_f32_dev(a.data_ptr(), rows, cols, v.data_ptr(), cols, o.data_ptr(), st)))
    t2 = timeit(lambda: trn.check(L.trn_vecmat_f32_dev(w.data_ptr(), rows, a.data_ptr(), rows, cols, o2.data_ptr(), st)))
    print(f"matvec {rows}x{cols}: {4.0*rows*cols/t/1e6:6.0f} GB/s   vecmat: {4.0*rows*cols/t2/1e6:6.0f} GB/s")
    del a
