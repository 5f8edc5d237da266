"""Epilogue timeline of one CTA of the pair GEMM (debug build with TRN_GEMM_TRACE) on BASELINE config 3."""
import os, sys
sys.path.insert(0, ".")
import torch
buf = torch.zeros(8 * 48, dtype=torch.int64, device="cuda")
os.environ["TRN_GEMM_TRACE"] = str(buf.data_ptr())
import trueno_b200 as trn
L = trn.lib
trn.check(L.trn_cuda_init(0))
B, H, m, k, n = 8, 32, 2048, 128, 2048
a = torch.rand(B * H * m * k, device="cuda"); b = torch.rand(B * H * k * n, device="cuda"); c = torch.empty(B * H * m * n, device="cuda")
trn.set_gemm_engine(2)
for _ in range(2):
    trn.check(L.trn_batched_matmul_4d_f32_dev(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), c.data_ptr(), B, H, m, k, n, None))
torch.cuda.synchronize()
t = buf.cpu().view(-1, 8)
t0 = int(t[0, 0])
print("tile: start  acc_drained  round0..3 (staging free)  stores_issued   [cycles]")
for i in range(4, 24):
    r = [int(x) - t0 for x in t[i][:7]]
    print(f"{i:2d}: {r[0]:7d} {r[1]:7d} | {r[2]:7d} {r[3]:7d} {r[4]:7d} {r[5]:7d} | {r[6]:7d}   period {int(t[i,0]) - int(t[i-1,0])}")
