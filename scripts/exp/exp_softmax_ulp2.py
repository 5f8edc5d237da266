"""Worst element error (ulps of the value) of every softmax row-kernel family at ~130 M elements per shape."""
import sys
sys.path.insert(0, ".")
import torch
import trueno_b200 as trn
torch.cuda.set_device(0); trn.check(trn.lib.trn_cuda_init(0))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
L = trn.lib
for rows, cols in [(131072, 1024), (262144, 512), (65536, 2048), (16384, 8192), (8192, 16384), (6000, 20000), (4096, 32000), (2600, 50257), (2048, 65536), (1024, 128256), (512, 262144), (4099, 32001), (16387, 8191)]:
    g = torch.Generator(device="cuda"); g.manual_seed(rows + cols)
    x = torch.randn(rows, cols, device="cuda", generator=g) * 4
    y = torch.empty_like(x)
    trn.check(L.trn_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, st)); torch.cuda.synchronize()
    worst = 0.0; wabs = 0.0
    step = max(1, (1 << 24) // cols)
    for r0 in range(0, rows, step):
        xs = x[r0:r0 + step]
        arg = (xs - xs.max(1, keepdim=True).values).double()
        e = torch.exp(arg)
        truth = e / e.sum(1, keepdim=True)
        ulp = 2.0 ** (torch.floor(torch.log2(truth.float().double().clamp_min(1e-300))) - 23)
        d = (y[r0:r0 + step].double() - truth).abs()
        worst = max(worst, (d / ulp).max().item()); wabs = max(wabs, d.max().item())
    print(f"{rows} x {cols}: worst {worst:.2f} ulp of the value, max abs {wabs:.2e}", flush=True)
    del x, y
