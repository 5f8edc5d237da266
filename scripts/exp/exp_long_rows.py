"""Long / unaligned softmax rows: the two-pass long kernel by cluster size and occupancy cap against the default dispatch.
TRN_ROWS_LONG_CS / TRN_ROWS_LONG_OCC are read per call by csrc/softmax.cu."""
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
for cols in [16387, 32001, 32768, 40000, 50257, 65536, 100003, 128256, 151936, 200019, 262144, 524288, 1 << 20]:
    rows = max(1, target // cols)
    x = torch.randn(rows, cols, device="cuda"); y = torch.empty_like(x)
    nb = 8.0 * rows * cols
    line = f"{rows:6d} x {cols:8d}:"
    for cs, occ in [(0, 0), (1, 0), (2, 0), (4, 0), (8, 0), (8, 4), (8, 2), (4, 4), (4, 2), (2, 2)]:
        if cs: os.environ["TRN_ROWS_LONG_CS"] = str(cs)
        else: os.environ.pop("TRN_ROWS_LONG_CS", None)
        os.environ["TRN_ROWS_LONG_OCC"] = str(occ)
        t1 = timeit(lambda: trn.check(L.trn_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
        t2 = timeit(lambda: trn.check(L.trn_log_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
        line += f"  [{'dflt' if not cs else f'cs{cs}/o{occ}'}] {nb/t1/1e6:5.0f}/{nb/t2/1e6:5.0f}"
    print(line, flush=True)
    del x, y
