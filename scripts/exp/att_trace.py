"""Timeline of one CTA of the attention kernel (debug build with TRN_ATT_TRACE): softmax-warp timestamps per key tile."""
import os, sys
sys.path.insert(0, ".")
import torch
buf = torch.zeros(1024, dtype=torch.int64, device="cuda")
os.environ["TRN_ATT_TRACE"] = str(buf.data_ptr())
import trueno_b200 as trn
L = trn.lib
trn.check(L.trn_cuda_init(0))
H, seq, d = 256, 2048, 128
q = torch.randn(H * seq * d, device="cuda"); k = torch.randn(H * seq * d, device="cuda"); v = torch.randn(H * seq * d, device="cuda")
o = torch.empty_like(q)
for _ in range(3):
    trn.check(L.trn_attention_f32_dev(q.data_ptr(), q.numel(), k.data_ptr(), k.numel(), v.data_ptr(), v.numel(), o.data_ptr(), H, seq, d, 1.0 / d ** 0.5, 0, None))
torch.cuda.synchronize()
t = buf.cpu()[:320].view(-1, 8)
t2 = buf.cpu()[512:832].view(-1, 8)
t0 = int(t[0, 0])
print("tile: wait_s  s_full  s_freed  p_computed  drained(prev)  p_stored | o_full(t) wake, S(t+1) issued  [cycles since CTA's first wait]")
for i in range(16):
    r = [int(x) - t0 if int(x) else 0 for x in t[i]]
    print(f"{i:2d}: {r[0]:7d} {r[1]:7d} {r[2]:7d} {r[3]:7d} {r[4]:7d} {r[5]:7d} | {r[6]:7d} {r[7]:7d}")
print("end", int(t[16, 0]) - t0)
print("drain(t): o_full wake -> ld0 done, fma0 done, ld1 done, fma1 done, arrived")
for i in range(15):
    w = int(t[i, 6]); print(i, [int(x) - w for x in t2[i][:5]])
