"""CTA-pair ring kernel (TRN_RING2=1) against the single-CTA ring: correctness on a spread of shapes, then microseconds per
call at the row counts a GPU owns when config 5 is sharded over 1/2/4/8 GPUs."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import trueno_b200 as trn
from trueno_b200 import parallel as par

torch.cuda.set_device(0); trn.check(trn.lib.trn_cuda_init(0))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
L = trn.lib

def run(x, log):
    y = torch.empty_like(x)
    f = L.trn_log_softmax_rows_f32_dev if log else L.trn_softmax_rows_f32_dev
    trn.check(f(x.data_ptr(), y.data_ptr(), x.shape[0], x.shape[1], st))
    torch.cuda.synchronize()
    return y

bad = 0
for rows, cols in ((1, 32000), (2, 32768), (3, 28680), (5, 30008), (147, 32000), (149, 32000), (512, 32000), (513, 31992), (1000, 32760)):
    g = torch.Generator(device="cuda"); g.manual_seed(rows * 7 + cols)
    x = torch.randn(rows, cols, device="cuda", generator=g) * 5
    if rows >= 3:
        x[1, : cols // 2] = -float("inf")          # one half of a row empty
        x[2, 7] = 80.0                               # a dominant element in the first half
    for log in (False, True):
        os.environ["TRN_RING2"] = "1"; y2 = run(x, log)
        os.environ["TRN_RING2"] = "0"; y1 = run(x, log)
        ref = (torch.log_softmax if log else torch.softmax)(x.double(), dim=1)
        if log:
            fin = torch.isfinite(ref)
            e2 = (y2.double() - ref)[fin].abs().max().item(); e1 = (y1.double() - ref)[fin].abs().max().item()
            same_inf = bool(((y2 == -float("inf")) == (ref == -float("inf"))).all())
        else:
            e2 = ((y2.double() - ref).abs() / ref.clamp_min(1e-30)).max().item(); e1 = ((y1.double() - ref).abs() / ref.clamp_min(1e-30)).max().item()
            same_inf = bool(((y2 == 0) == (y1 == 0)).all())
        ok = e2 < (2e-6 if log else 4e-6) * (1 if not log else 10) and same_inf and not torch.isnan(y2).any()
        bad += not ok
        print(f"{rows}x{cols} log={log}: ring2 err {e2:.3e} ring err {e1:.3e} {'ok' if ok else 'BAD'}", flush=True)
print("correctness:", "PASS" if not bad else f"{bad} BAD", flush=True)

def timeit(fn, iters=20):
    loop = par.CapturedLoop(fn, iters)
    loop.replay(); loop.replay(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        loop.replay(); s.record(stream); loop.replay(); e.record(stream); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / iters)
    return best * 1e3

cols = 32000
for rows in (512, 1024, 2048, 4096, 8192):
    x = torch.randn(rows, cols, device="cuda") * 4; y = torch.empty_like(x)
    line = f"{rows:5d} rows (ideal at 6.45 TB/s {8.0 * rows * cols / 6.4549e6:6.1f} us):"
    for r2 in ("0", "1"):
        os.environ["TRN_RING2"] = r2
        t1 = timeit(lambda: trn.check(L.trn_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
        t2 = timeit(lambda: trn.check(L.trn_log_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)))
        line += f"  [ring2={r2}] {t1:6.1f}/{t2:6.1f} us = {8e-6 * rows * cols / t1:5.2f}/{8e-6 * rows * cols / t2:5.2f} TB/s"
    print(line, flush=True)
    del x, y
