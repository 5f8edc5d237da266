"""config-5 rows (32 000 columns) at the row counts a GPU owns when 4096 rows are sharded over 1/2/4/8 GPUs:
ring kernel (16 / 32 KiB slots) vs the two-pass cluster kernel, microseconds per call and the ideal at 6.2 TB/s."""
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
for rows in (512, 1024, 2048, 4096, 8192):
    x = torch.randn(rows, cols, device="cuda") * 4; y = torch.empty_like(x)
    line = f"{rows:5d} rows (ideal {8.0 * rows * cols / 6.2e6:6.1f} us):"
    for env in ({}, {"TRN_RING_HPC": "2"}, {"TRN_RING_HPC": "4"}, {"TRN_ROWS_LONG_CS": "1"}, {"TRN_ROWS_LONG_CS": "2"}, {"TRN_ROWS_LONG_CS": "4"}):
        for k in ("TRN_RING_HPC", "TRN_ROWS_LONG_CS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        t1 = timeit(lambda: trn.check(L.trn_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
        t2 = timeit(lambda: trn.check(L.trn_log_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
        tag = "default" if not env else "ring%s" % (int(env["TRN_RING_HPC"]) * 8) if "TRN_RING_HPC" in env else "2pass-cs" + env["TRN_ROWS_LONG_CS"]
        line += f"  [{tag}] {t1:6.1f}/{t2:6.1f}"
    print(line, flush=True)
    del x, y
# maps at the same sizes
for n in (16_384_000, 32_768_000, 65_536_000, 131_072_000):
    x = torch.randn(n, device="cuda"); y = torch.empty_like(x)
    t = timeit(lambda: trn.check(L.trn_gelu_f32_dev(x.data_ptr(), n, y.data_ptr(), st)))
    t2 = timeit(lambda: trn.check(L.trn_add_f32_dev(x.data_ptr(), n, x.data_ptr(), n, y.data_ptr(), st)))
    print(f"gelu {n}: {t:6.1f} us (ideal {8.0 * n / 6.6e6:6.1f})   add: {t2:6.1f} us (ideal {12.0 * n / 6.8e6:6.1f})")
