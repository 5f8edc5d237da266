import os, sys
import torch
sys.path.insert(0, ".")
import trueno_b200 as trn
def timeit(fn, iters=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters
torch.cuda.set_device(0); trn.check(trn.lib.trn_cuda_init(0))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
L = trn.lib
target = 1 << 27
for cols in [16388, 17000, 18432, 20000, 33000, 36000, 40000, 45000, 50257, 57344, 65536, 80000, 98304, 100003, 128256]:
    rows = target // cols
    x = torch.randn(rows, cols, device="cuda") * 4; y = torch.empty_like(x)
    nb = 8.0 * rows * cols
    line = f"{rows:6d} x {cols:6d}:"
    for cs in (1, 2, 4, 8):
        os.environ["TRN_ROWS_LONG_CS"] = str(cs)
        t1 = timeit(lambda: trn.check(L.trn_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
        t2 = timeit(lambda: trn.check(L.trn_log_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
        line += f"  [cs{cs}] {nb/t1/1e6:5.0f}/{nb/t2/1e6:5.0f}"
    print(line, flush=True)
    del x, y
