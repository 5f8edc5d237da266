import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import trueno_b200 as trn
from trueno_b200 import parallel as par
torch.cuda.set_device(0); trn.check(trn.lib.trn_cuda_init(0))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
L = trn.lib
def timeit(fn, iters=20):
    loop = par.CapturedLoop(fn, iters)
    loop.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream); loop.replay(); e.record(stream); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / iters)
    return best * 1e3
cols = 32000
for rows in (512, 1024, 2048):
    x = torch.randn(rows, cols, device="cuda") * 4; y = torch.empty_like(x)
    line = f"{rows} rows:"
    for cs in ("0", "1", "2", "4", "8"):
        os.environ.pop("TRN_ROWS_LONG_CS", None)
        if cs != "0": os.environ["TRN_ROWS_LONG_CS"] = cs
        t1 = timeit(lambda: trn.check(L.trn_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
        t1s = timeit(lambda: trn.check(L.trn_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)), iters=1)
        line += f"  [cs{cs}] {t1:5.1f} (single {t1s:5.1f})"
    print(line, flush=True)
