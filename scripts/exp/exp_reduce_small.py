"""One GPU's slice of config 4 (2^27 f32 at 8 GPUs, 2^28 at 4, 2^29 at 2): microseconds per call inside a CUDA graph and the
streaming ideal, by resident blocks per SM (TRN_REDUCE_PER_SM, read once per process)."""
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
line = f"per_sm={os.environ.get('TRN_REDUCE_PER_SM', 'max')}:"
for lg in (27, 28, 30):
    n = 1 << lg
    x = torch.rand(n, device="cuda") * 2 - 1
    t_sum = timeit(lambda: trn.check(L.trn_sum_f32_dev(x.data_ptr(), n, out.data_ptr(), st)))
    t_arg = timeit(lambda: trn.check(L.trn_argmax_f32_dev(x.data_ptr(), n, oi.data_ptr(), out.data_ptr(), st)))
    t_mv = 0
    line += f"  2^{lg}: sum {t_sum:6.1f} argmax {t_arg:6.1f} (ideal {4.0 * n / 7.2e6:6.1f})"
    del x
print(line)
