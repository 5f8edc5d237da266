"""gelu variants: accuracy against the scalar-backend oracle's formula in f64 and microseconds at config-5 size."""
import os, sys
sys.path.insert(0, ".")
import torch
import trueno_b200 as trn
from trueno_b200 import parallel as par
torch.cuda.set_device(0); trn.check(trn.lib.trn_cuda_init(0))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
L = trn.lib
n = 131_072_000
for name, x in (("N(0,1)*4", torch.randn(n, device="cuda") * 4), ("U[-12,12]", torch.rand(n, device="cuda") * 24 - 12), ("U[-3,3]", torch.rand(n, device="cuda") * 6 - 3)):
    y = torch.empty_like(x)
    trn.check(L.trn_gelu_f32_dev(x.data_ptr(), n, y.data_ptr(), st)); torch.cuda.synchronize()
    # the reference's expression with its f32 u (src/backends/scalar.rs:330-340), the rest in f64
    x3 = (x * x) * x
    u = (0.7978846 * torch.ones((), device="cuda", dtype=torch.float32)) * (x + (0.044715 * torch.ones((), device="cuda", dtype=torch.float32)) * x3)
    truth = x.double() / (1.0 + torch.exp(-2.0 * u.double()))
    err = (y.double() - truth).abs()
    ulp = torch.maximum(torch.abs(truth).float(), torch.tensor(1e-45, device="cuda")).double()
    ulp_y = 2.0 ** (torch.floor(torch.log2(ulp)) - 23)
    bound = 4 * ulp_y + 4 * 2.0 ** -24 * x.double().abs()
    print(f"{name}: max err / ulp(y) = {(err / ulp_y).max().item():.2f}   max err / bound = {(err / bound).max().item():.3f}   max abs {err.max().item():.2e}", flush=True)
    del x3, u, truth, err, ulp, ulp_y, bound
x = torch.randn(n, device="cuda") * 4; y = torch.empty_like(x)
loop = par.CapturedLoop(lambda: trn.check(L.trn_gelu_f32_dev(x.data_ptr(), n, y.data_ptr(), st)), 20)
loop2 = par.CapturedLoop(lambda: trn.check(L.trn_sigmoid_f32_dev(x.data_ptr(), n, y.data_ptr(), st)), 20)
for lp, nm in ((loop, "gelu"), (loop2, "sigmoid")):
    ts = []
    for _ in range(6):
        s_, e_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lp.replay(); s_.record(stream); lp.replay(); e_.record(stream); torch.cuda.synchronize()
        ts.append(s_.elapsed_time(e_) / 20 * 1e3)
    print(f"{nm}: min {min(ts):.1f} us  median {sorted(ts)[3]:.1f} us  = {8e-6 * n / min(ts):.2f} TB/s")
