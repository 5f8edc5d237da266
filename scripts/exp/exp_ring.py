"""Rows of 16384 < cols <= 32768: the TMA ring kernel by slot size (TRN_RING_HPC = 2 / 4: 16 / 32 KiB) against the
two-pass cluster kernel (TRN_ROWS_LONG_CS = 1 / 2 / 4); both knobs are read per call."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
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
for rows, cols in [(7895, 17000), (6710, 20000), (5461, 24576), (4793, 28000), (4096, 32000), (4194, 32001), (4096, 32768), (512, 32000), (1024, 32000), (2048, 32000), (3072, 32000), (3552, 32000), (8192, 32000)]:
    x = torch.randn(rows, cols, device="cuda") * 4; y = torch.empty_like(x)
    nb = 8.0 * rows * cols
    line = f"{rows:6d} x {cols:6d}:"
    for env in ({}, {"TRN_RING_HPC": "2"}, {"TRN_RING_HPC": "4"}, {"TRN_ROWS_LONG_CS": "1"}, {"TRN_ROWS_LONG_CS": "2"}, {"TRN_ROWS_LONG_CS": "4"}):
        for k in ("TRN_RING_HPC", "TRN_ROWS_LONG_CS"):
            os.environ.pop(k, None)
        os.environ.update(env)
        t1 = timeit(lambda: trn.check(L.trn_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
        t2 = timeit(lambda: trn.check(L.trn_log_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
        tag = "default" if not env else "ring %s KiB" % (int(env["TRN_RING_HPC"]) * 8) if "TRN_RING_HPC" in env else "two-pass cs" + env["TRN_ROWS_LONG_CS"]
        line += f"  [{tag}] {nb/t1/1e6:5.0f}/{nb/t2/1e6:5.0f}"
    print(line, flush=True)
    del x, y
