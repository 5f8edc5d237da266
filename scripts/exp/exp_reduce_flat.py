"""Flat-grid reductions (TRN_REDUCE_FLAT = 16 KiB tiles per block, read per call; 0 = one persistent wave): microseconds per
call inside a CUDA graph at one GPU's slice of config 4 over 8 / 4 / 1 GPUs, with the result checked against f64."""
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
    for _ in range(3): loop.replay()
    best = 1e9
    for _ in range(3):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(stream); loop.replay(); loop.replay(); e.record(stream); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e) / (2 * iters))
    return best * 1e3
out = torch.zeros(4, device="cuda"); oi = torch.zeros(2, dtype=torch.int64, device="cuda")
for lg in (27, 28, 30):
    n = 1 << lg
    x = torch.rand(n, device="cuda") * 2 - 1
    x[n - 12345] = 3.0
    truth = x.double().sum().item(); scale = x.double().abs().sum().item()
    for flat in (0, 1, 2, 4, 8, 16, 32):
        os.environ["TRN_REDUCE_FLAT"] = str(flat)
        t_sum = timeit(lambda: trn.check(L.trn_sum_f32_dev(x.data_ptr(), n, out.data_ptr(), st)))
        torch.cuda.synchronize(); err = abs(out[0].item() - truth) / scale
        t_arg = timeit(lambda: trn.check(L.trn_argmax_f32_dev(x.data_ptr(), n, oi.data_ptr(), out.data_ptr(), st)))
        torch.cuda.synchronize(); ok = oi[0].item() == n - 12345
        t_dot = timeit(lambda: trn.check(L.trn_dot_f32_dev(x.data_ptr(), n, x.data_ptr(), n, out.data_ptr(), st))) if lg < 30 else 0
        print(f"2^{lg} flat={flat:2d}: sum {t_sum:6.1f} us (err {err:.1e})  argmax {t_arg:6.1f} us ({'ok' if ok else 'WRONG'})  dot(x,x) {t_dot:6.1f}   ideal at 7.2 TB/s {4.0 * n / 7.2e6:6.1f}", flush=True)
    del x
