"""Pure-write bandwidth (SM -> L2 -> HBM) of plain vectorised stores: torch fill_ on 4 GiB, and cudaMemset."""
import torch
x = torch.empty(1 << 30, device="cuda")
for name, fn in (("fill_", lambda: x.fill_(1.0)), ("zero_ (memset)", lambda: x.zero_())):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
    for s, e in evs:
        s.record(); fn(); e.record()
    torch.cuda.synchronize()
    ms = sorted(s.elapsed_time(e) for s, e in evs)[5]
    print(f"{name}: {ms:.3f} ms  {4 * x.numel() / ms / 1e6:.0f} GB/s written")
y = torch.empty(1 << 22, device="cuda")   # 16 MiB: stays in L2
for _ in range(3): y.fill_(1.0)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(200): y.fill_(1.0)
e.record(); torch.cuda.synchronize()
print(f"fill_ 16 MiB x200 (L2-resident): {4 * y.numel() * 200 / s.elapsed_time(e) / 1e6:.0f} GB/s written")
