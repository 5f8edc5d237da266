"""Vector-length sweep: reductions and maps from 2^16 to 2^30 elements (GB/s and microseconds, device-resident)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trueno_b200 as trn

def timeit(fn, iters=50):
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
out = torch.zeros(4, device="cuda"); oi = torch.zeros(2, dtype=torch.int64, device="cuda")
for p in (16, 18, 20, 22, 24, 26, 28, 30):
    n = 1 << p
    a = torch.rand(n, device="cuda") * 2 - 1; b = torch.rand(n, device="cuda"); o = torch.empty(n, device="cuda")
    it = 50 if p < 28 else 20
    ts = timeit(lambda: trn.check(L.trn_sum_f32_dev(a.data_ptr(), n, out.data_ptr(), st)), it)
    td = timeit(lambda: trn.check(L.trn_dot_f32_dev(a.data_ptr(), n, b.data_ptr(), n, out.data_ptr(), st)), it)
    ta = timeit(lambda: trn.check(L.trn_argmax_f32_dev(a.data_ptr(), n, oi.data_ptr(), out.data_ptr(), st)), it)
    tm = timeit(lambda: trn.check(L.trn_add_f32_dev(a.data_ptr(), n, b.data_ptr(), n, o.data_ptr(), st)), it)
    tg = timeit(lambda: trn.check(L.trn_gelu_f32_dev(a.data_ptr(), n, o.data_ptr(), st)), it)
    tt = timeit(lambda: torch.sum(a), it)
    print(f"n=2^{p}: sum {ts*1e3:8.1f} us {4*n/ts/1e6:6.0f} GB/s | dot {td*1e3:8.1f} us {8*n/td/1e6:6.0f} | argmax {ta*1e3:8.1f} us {4*n/ta/1e6:6.0f} | add {tm*1e3:8.1f} us {12*n/tm/1e6:6.0f} | gelu {tg*1e3:8.1f} us {8*n/tg/1e6:6.0f} | torch.sum {tt*1e3:8.1f} us")
    del a, b, o
