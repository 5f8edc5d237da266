"""Worst element error of the ring softmax kernel, in ulps of the value, on the inputs of
tests/test_parity_gpu.py::test_ring_kernel_claimed_rows_equal_dealt_rows (same generator, same order) and on config 5."""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
import trueno_b200 as trn
torch.cuda.set_device(0); trn.check(trn.lib.trn_cuda_init(0))
rng = np.random.default_rng(77)
f32 = np.float32
shapes = [(1, 32000), (3, 28680), (149, 32000), (513, 31992), (1200, 32768)]
extra = [(4096, 32000)]
for rows, cols in shapes + extra:
    x = (rng.standard_normal((rows, cols)) * 4).astype(f32) if (rows, cols) in shapes else (np.random.default_rng(5).standard_normal((rows, cols)) * 4).astype(f32)
    got = torch.from_numpy(trn.softmax_rows(x, rows, cols)).cuda().double()
    xt = torch.from_numpy(x).cuda()
    arg = (xt - xt.max(1, keepdim=True).values).double()          # the f32 subtraction, exactly representable in f64
    e = torch.exp(arg)
    truth = e / e.sum(1, keepdim=True)
    ulp = 2.0 ** (torch.floor(torch.log2(truth.float().double().clamp_min(1e-300))) - 23)
    r = ((got - truth).abs() / ulp)
    print(f"{rows} x {cols}: worst {r.max().item():.2f} ulp of the value, 99.999th percentile {torch.quantile(r.flatten()[:16_000_000].float(), 0.99999).item():.2f}, max abs {(got - truth).abs().max().item():.2e}", flush=True)
