"""Row kernels beyond the aligned <= 65536-column cases: window form (cols % 4 != 0), long rows (LLM vocabularies),
few-long-rows split form (Vector::softmax on one large vector), layer_norm up to 16384.  GB/s = 8 B per element / time.
Run once as is and once with TRN_ROWS_GENERIC=1 (the three-pass fallback these kernels replace)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trueno_b200 as trn

def timeit(fn, iters=10):
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
tag = "generic" if os.environ.get("TRN_ROWS_GENERIC") == "1" else "new"
with_torch = tag == "new"
target = 1 << 27   # elements per shape (512 MiB in, 512 MiB out)
shapes = [77, 1001, 4099, 8191, 16387, 32001, 50257, 65536, 65540, 100003, 128256, 151936, 262144, 1 << 20]
for cols in shapes:
    rows = max(1, target // cols)
    x = torch.randn(rows, cols, device="cuda"); y = torch.empty_like(x)
    nb = 8.0 * rows * cols
    t1 = timeit(lambda: trn.check(L.trn_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
    t2 = timeit(lambda: trn.check(L.trn_log_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
    t5 = timeit(lambda: torch.softmax(x, dim=1, out=y)) if with_torch else float("nan")
    print(f"[{tag}] {rows:8d} x {cols:8d}: softmax {nb/t1/1e6:6.0f}  log_softmax {nb/t2/1e6:6.0f}  torch.softmax {nb/t5/1e6:6.0f} GB/s", flush=True)
    del x, y
# few long rows: one Vector::softmax
for rows, cols in [(1, 1 << 20), (1, 1 << 24), (1, (1 << 27) + 1), (1, 1 << 28), (4, 1 << 24), (16, 1 << 22)]:
    x = torch.randn(rows, cols, device="cuda"); y = torch.empty_like(x)
    nb = 8.0 * rows * cols
    t1 = timeit(lambda: trn.check(L.trn_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)), iters=5)
    t2 = timeit(lambda: trn.check(L.trn_log_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)), iters=5)
    t5 = timeit(lambda: torch.softmax(x, dim=1, out=y), iters=5) if with_torch else float("nan")
    print(f"[{tag}] {rows:8d} x {cols:9d}: softmax {nb/t1/1e6:6.0f}  log_softmax {nb/t2/1e6:6.0f}  torch.softmax {nb/t5/1e6:6.0f} GB/s  ({t1*1e3:.0f} us)", flush=True)
    del x, y
if tag == "new":
    for rows, cols in [(1 << 14, 8192), (10922, 12288), (8192, 16384), (6553, 20480), (4096, 32768), (2048, 65536), (1024, 131072), (512, 262144)]:
        x = torch.randn(rows, cols, device="cuda"); y = torch.empty_like(x)
        g = torch.randn(cols, device="cuda"); b = torch.randn(cols, device="cuda")
        nb = 8.0 * rows * cols
        t3 = timeit(lambda: trn.check(L.trn_layer_norm_rows_f32_dev(x.data_ptr(), g.data_ptr(), cols, b.data_ptr(), cols, 1e-5, y.data_ptr(), rows, cols, st)))
        t6 = timeit(lambda: torch.nn.functional.layer_norm(x, (cols,), g, b, 1e-5))
        print(f"[{tag}] {rows:8d} x {cols:8d}: layer_norm {nb/t3/1e6:6.0f}  torch.layer_norm {nb/t6/1e6:6.0f} GB/s", flush=True)
        del x, y
if tag == "new":
    # Matrix::embedding_lookup: 8 B per output element (4 read + 4 written)
    for rows, cols, n in [(50257, 768, 1 << 19), (32000, 4096, 1 << 16), (128256, 8192, 1 << 14), (1000, 64, 1 << 22), (50257, 1001, 1 << 18)]:
        table = torch.randn(rows, cols, device="cuda")
        idx = torch.randint(0, rows, (n,), device="cuda", dtype=torch.int64)
        out = torch.empty(n, cols, device="cuda")
        nb = 8.0 * n * cols
        t1 = timeit(lambda: trn.check(L.trn_embedding_lookup_f32_dev(table.data_ptr(), rows, cols, idx.data_ptr(), n, out.data_ptr(), st)))
        t2 = timeit(lambda: torch.index_select(table, 0, idx, out=out))
        print(f"[{tag}] embedding_lookup {rows} x {cols}, {n} indices: {nb/t1/1e6:6.0f}  torch.index_select {nb/t2/1e6:6.0f} GB/s", flush=True)
        del table, out
