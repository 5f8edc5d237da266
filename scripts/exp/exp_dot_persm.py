"""dot 2^30 by resident CTAs per SM (TRN_REDUCE_PER_SM, read once per process)."""
import os, sys
sys.path.insert(0, ".")
import torch
import trueno_b200 as trn
from trueno_b200 import parallel as par
L = trn.lib
torch.cuda.set_device(0); trn.check(L.trn_cuda_init(0))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
out = torch.zeros(1, device="cuda")
line = f"per_sm={os.environ.get('TRN_REDUCE_PER_SM', 'default(4)')}:"
for lg in (27, 30):
    n = 1 << lg
    a = torch.rand(n, device="cuda"); b = torch.rand(n, device="cuda")
    lp = par.CapturedLoop(lambda: trn.check(L.trn_dot_f32_dev(a.data_ptr(), n, b.data_ptr(), n, out.data_ptr(), st)), 10)
    ts = []
    for _ in range(6):
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lp.replay(); s_.record(stream); lp.replay(); e_.record(stream); torch.cuda.synchronize()
        ts.append(s_.elapsed_time(e_) / 10 * 1e3)
    line += f"  2^{lg}: {min(ts):.1f} us = {8e-6 * n / min(ts):.3f} TB/s"
    del a, b, lp
print(line)
