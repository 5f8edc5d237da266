"""Config 3 timing for A/B runs across processes (TRN_GEMM_DEBUG etc. are read once per process): min / median of 40 timed
pairs of calls, each after a 2 ms pause (burst regime, as bench.py sees it), on the SURVEY generator's data."""
import os, sys, time, statistics
sys.path.insert(0, ".")
import torch
import trueno_b200 as trn
L = trn.lib
torch.cuda.set_device(0); trn.check(L.trn_cuda_init(0))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
B, H, m, k, n = 8, 32, 2048, 128, 2048
a = torch.rand(B * H * m * k, device="cuda"); b = torch.rand(B * H * k * n, device="cuda"); c = torch.empty(B * H * m * n, device="cuda")
f = lambda: trn.check(L.trn_batched_matmul_4d_f32_dev(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), c.data_ptr(), B, H, m, k, n, st))
for _ in range(5): f()
torch.cuda.synchronize()
ts = []
for _ in range(40):
    time.sleep(0.002)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f(); e0.record(stream); f(); f(); e1.record(stream); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 2)
chk = float(c[:: 1 << 20].double().sum())
print(f"TRN_GEMM_DEBUG={os.environ.get('TRN_GEMM_DEBUG', '0')}: min {min(ts):.4f} median {statistics.median(ts):.4f} ms  checksum {chk:.6f}")
