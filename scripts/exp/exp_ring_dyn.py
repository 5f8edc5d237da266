"""Ring softmax kernel with CLAIMED rows (TRN_RING_DYN=1, default) against dealt rows (=0): correctness on a spread of shapes
(also after hundreds of back-to-back launches: the claim counters must come back to zero), then microseconds per call at the
row counts a GPU owns when config 5 is sharded over 1/2/4/8 GPUs."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import trueno_b200 as trn
from trueno_b200 import parallel as par

torch.cuda.set_device(0); trn.check(trn.lib.trn_cuda_init(0))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
L = trn.lib

def run(x, log):
    y = torch.full_like(x, float("nan"))
    f = L.trn_log_softmax_rows_f32_dev if log else L.trn_softmax_rows_f32_dev
    trn.check(f(x.data_ptr(), y.data_ptr(), x.shape[0], x.shape[1], st))
    torch.cuda.synchronize()
    return y

def check(tag):
    bad = 0
    for rows, cols in ((1, 32000), (2, 32768), (3, 28680), (147, 32000), (149, 32000), (512, 32000), (513, 31992), (1000, 32760), (4096, 32000)):
        g = torch.Generator(device="cuda"); g.manual_seed(rows * 7 + cols)
        x = torch.randn(rows, cols, device="cuda", generator=g) * 5
        for log in (False, True):
            os.environ["TRN_RING_DYN"] = "1"; y1 = run(x, log)
            os.environ["TRN_RING_DYN"] = "0"; y0 = run(x, log)
            ok = bool(torch.equal(y1, y0)) and not bool(torch.isnan(y1).any())
            bad += not ok
            if not ok:
                print(f"{tag}: {rows}x{cols} log={log}: MISMATCH ({int((y1 != y0).sum())} elements, {int(torch.isnan(y1).sum())} NaN)", flush=True)
    print(f"{tag}: claimed rows == dealt rows bit for bit:", "PASS" if not bad else f"{bad} BAD", flush=True)

check("cold")

def timeit(fn, iters=20):
    loop = par.CapturedLoop(fn, iters)
    loop.replay(); loop.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        loop.replay(); s.record(stream); loop.replay(); e.record(stream); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / iters)
    return best * 1e3

import statistics
cols = 32000
for rows in (512, 1024, 2048, 4096, 8192):
    x = torch.randn(rows, cols, device="cuda") * 4; y = torch.empty_like(x)
    loops = {}
    for dyn in ("0", "1"):
        os.environ["TRN_RING_DYN"] = dyn           # read at capture time
        loops[dyn, 0] = par.CapturedLoop(lambda: trn.check(L.trn_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)), 20)
        loops[dyn, 1] = par.CapturedLoop(lambda: trn.check(L.trn_log_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)), 20)
    ts = {k: [] for k in loops}
    for rep in range(12):                          # interleaved: drift of the box hits both arms alike
        for k, loop in loops.items():
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            loop.replay(); s_.record(stream); loop.replay(); e_.record(stream); torch.cuda.synchronize()
            ts[k].append(s_.elapsed_time(e_) / 20 * 1e3)
    line = f"{rows:5d} rows (ideal at 6.45 TB/s {8.0 * rows * cols / 6.4549e6:6.1f} us):"
    for log in (0, 1):
        for dyn in ("0", "1"):
            v = ts[dyn, log]
            line += f"  [{'log_' if log else ''}softmax dyn={dyn}] min {min(v):6.1f} med {statistics.median(v):6.1f}"
    print(line, flush=True)
    del x, y, loops
check("after the timing loops")
