"""Does the distance between the two operands of dot matter (channel camping between the a and b streams)?"""
import sys, statistics
sys.path.insert(0, ".")
import torch
import trueno_b200 as trn
from trueno_b200 import parallel as par
L = trn.lib
torch.cuda.set_device(0); trn.check(L.trn_cuda_init(0))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
n = 1 << 30
big = torch.empty(2 * n + (1 << 22), device="cuda")
big.uniform_(-1, 1)
out = torch.zeros(1, device="cuda")
loops = {}
for pad in (0, 256, 1024, 4096, 16384, 65536 + 1024, (1 << 20) + 4096):   # elements
    a, b = big[:n], big[n + pad: 2 * n + pad]
    loops[pad] = par.CapturedLoop(lambda a=a, b=b: trn.check(L.trn_dot_f32_dev(a.data_ptr(), n, b.data_ptr(), n, out.data_ptr(), st)), 10)
loops["sum"] = par.CapturedLoop(lambda: trn.check(L.trn_sum_f32_dev(big.data_ptr(), 2 * n, out.data_ptr(), st)), 10)
ts = {k: [] for k in loops}
for rep in range(5):
    for k, lp in loops.items():
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lp.replay(); s_.record(stream); lp.replay(); e_.record(stream); torch.cuda.synchronize()
        ts[k].append(s_.elapsed_time(e_) / 10 * 1e3)
for k in loops:
    print(f"pad {k}: min {min(ts[k]):.1f} us  {8e-6 * n / min(ts[k]):.3f} TB/s")
