"""Per-CTA end times of the A-stationary config-3 kernel (TRN_GEMM_TRACE=2): is the static unit deal ending on slow pairs?"""
import os, sys
os.environ["TRN_GEMM_TRACE"] = "2"
sys.path.insert(0, ".")
import torch
import trueno_b200 as trn
L = trn.lib
torch.cuda.set_device(0); trn.check(L.trn_cuda_init(0))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
B, H, m, k, n = 8, 32, 2048, 128, 2048
a = torch.rand(B * H * m * k, device="cuda"); b = torch.rand(B * H * k * n, device="cuda"); c = torch.empty(B * H * m * n, device="cuda")
for i in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    trn.check(L.trn_batched_matmul_4d_f32_dev(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), c.data_ptr(), B, H, m, k, n, st))
    e1.record(stream); torch.cuda.synchronize()
    print(f"call {i}: {e0.elapsed_time(e1):.3f} ms", flush=True)
