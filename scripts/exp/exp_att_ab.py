"""Fused attention timing / accuracy for A/B runs across processes (TRN_ATT_ROT is read once per process)."""
import os, sys, time, statistics
sys.path.insert(0, ".")
import torch
import trueno_b200 as trn
L = trn.lib
torch.cuda.set_device(0); trn.check(L.trn_cuda_init(0))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
H, S, D = 256, 2048, 128
g = torch.Generator(device="cuda"); g.manual_seed(3)
q, k, v = (torch.randn(H * S * D, device="cuda", generator=g) for _ in range(3))
o = torch.empty_like(q)
f = lambda: trn.check(L.trn_attention_f32_dev(q.data_ptr(), q.numel(), k.data_ptr(), k.numel(), v.data_ptr(), v.numel(), o.data_ptr(), H, S, D, 1.0 / D ** 0.5, 0, st))
for _ in range(3): f()
torch.cuda.synchronize()
ts = []
for _ in range(25):
    time.sleep(0.002)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f(); e0.record(stream); f(); f(); e1.record(stream); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) / 2)
# accuracy of head 5 against f64
h = 5
qq, kk, vv = (t.view(H, S, D)[h].double() for t in (q, k, v))
ref = torch.softmax(qq @ kk.T / D ** 0.5, dim=1) @ vv
err = (o.view(H, S, D)[h].double() - ref).abs().max().item()
print(f"TRN_ATT_RAW_HI={os.environ.get('TRN_ATT_RAW_HI', '1')}: min {min(ts):.4f} median {statistics.median(ts):.4f} ms  max abs err head 5 {err:.2e}")
