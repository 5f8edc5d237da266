"""Long / unaligned softmax rows: the two-pass long kernel by cluster size against the default dispatch.
TRN_ROWS_LONG_CS is read per call by csrc/softmax.cu.  (An occupancy cap — fewer rows in flight against L2 — and a
shared-memory-resident one-pass cluster kernel were measured here too and were slower at every length; both are gone.)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
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
target = 1 << 27
def run(cols, env):
    for k in ("TRN_ROWS_LONG_CS", "TRN_ROWS_L2HINT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    t1 = timeit(lambda: trn.check(L.trn_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
    t2 = timeit(lambda: trn.check(L.trn_log_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
    return f"{nb/t1/1e6:5.0f}/{nb/t2/1e6:5.0f}"

for cols in [32000, 40000, 50257, 65536, 65540, 100003, 128256, 131072, 151936, 196608, 200019, 262144, 524288, 1 << 20]:
    rows = max(1, target // cols)
    x = torch.randn(rows, cols, device="cuda"); y = torch.empty_like(x)
    nb = 8.0 * rows * cols
    line = f"{rows:6d} x {cols:8d}:  [default] {run(cols, {})}  [no L2 hints] {run(cols, {'TRN_ROWS_L2HINT': '1'})}"
    if cols > 32768:
        for cs in (4, 8):
            line += f"  [cs{cs}] {run(cols, {'TRN_ROWS_LONG_CS': str(cs)})}  [cs{cs}, no hints] {run(cols, {'TRN_ROWS_LONG_CS': str(cs), 'TRN_ROWS_L2HINT': '1'})}"
    print(line, flush=True)
    del x, y
