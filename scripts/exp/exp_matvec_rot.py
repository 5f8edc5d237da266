"""matvec 32768^2 (and one GPU's 4096-row share): rows walked from a row-dependent block on (TRN_MATVEC_ROT, read per call)."""
import os, sys, statistics
sys.path.insert(0, ".")
import torch
import trueno_b200 as trn
from trueno_b200 import parallel as par
L = trn.lib
torch.cuda.set_device(0); trn.check(L.trn_cuda_init(0))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
for rows, cols in ((32768, 32768), (4096, 32768), (16384, 16384), (8192, 65536)):
    a = torch.rand(rows, cols, device="cuda"); v = torch.rand(cols, device="cuda") * 2 - 1; y = torch.empty(rows, device="cuda")
    ref = None
    loops = {}
    for rot in ("0", "1", "3", "5"):
        os.environ["TRN_MATVEC_ROT"] = rot
        loops[rot] = par.CapturedLoop(lambda: trn.check(L.trn_matvec_f32_dev(a.data_ptr(), rows, cols, v.data_ptr(), cols, y.data_ptr(), st)), 10)
    ts = {k: [] for k in loops}
    for rep in range(8):
        for kk, lp in loops.items():
            s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            lp.replay(); s_.record(stream); lp.replay(); e_.record(stream); torch.cuda.synchronize()
            ts[kk].append(s_.elapsed_time(e_) / 10 * 1e3)
    line = f"{rows} x {cols}:"
    for kk in loops:
        line += f"  rot={kk}: min {min(ts[kk]):7.1f} med {statistics.median(ts[kk]):7.1f} us ({4e-6 * rows * cols / min(ts[kk]):.2f} TB/s)"
    print(line, flush=True)
    del a, v, y, loops
