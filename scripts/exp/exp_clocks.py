"""Sample SM clock / power / throttle reasons while one op loops (is a kernel power-bound?)."""
import sys, threading, time, statistics
sys.path.insert(0, ".")
import torch, pynvml
import trueno_b200 as trn
L = trn.lib
torch.cuda.set_device(0)
trn.check(L.trn_cuda_init(0))
_stream = torch.cuda.Stream()
torch.cuda.set_stream(_stream)
st = _stream.cuda_stream
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)

def sample(stop, out):
    while not stop.is_set():
        out.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                    pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
        time.sleep(0.02)

def run(name, fn, secs=3.0):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    stop, out = threading.Event(), []
    t = threading.Thread(target=sample, args=(stop, out)); t.start()
    t0 = time.time(); n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < secs:
        for _ in range(20): fn()
        n += 20
        torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    stop.set(); t.join()
    out = out[len(out) // 3:]
    reasons = 0
    for o in out: reasons |= o[2]
    print(f"{name}: {e0.elapsed_time(e1) / n:.3f} ms/op  sm_mhz median {statistics.median(o[0] for o in out)} min {min(o[0] for o in out)} "
          f"power median {statistics.median(o[1] for o in out):.0f} W max {max(o[1] for o in out):.0f} W reasons 0x{reasons:x}")

trn.set_gemm_engine(2)
B, H, m, k, n = 8, 32, 2048, 128, 2048
import os
for kind in ("U[0,1)", "N(0,1)"):
    a = torch.rand(B * H * m * k, device="cuda") if kind[0] == "U" else torch.randn(B * H * m * k, device="cuda")
    b = torch.rand(B * H * k * n, device="cuda") if kind[0] == "U" else torch.randn(B * H * k * n, device="cuda")
    c = torch.empty(B * H * m * n, device="cuda")
    f = lambda: trn.check(L.trn_batched_matmul_4d_f32_dev(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), c.data_ptr(), B, H, m, k, n, st))
    run(f"batched4d {kind} sustained 3 s", f)
    # burst: one call at a time from an idle-ish GPU (2 ms pause), best and median of 30
    ts = []
    for _ in range(30):
        time.sleep(0.002)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f(); e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    print(f"batched4d {kind} burst (pairs of calls after a 2 ms pause, 2nd timed): best {min(ts):.3f} median {statistics.median(ts):.3f} ms")
    del a, b, c
N = 8192
a = torch.rand(N * N, device="cuda"); b = torch.rand(N * N, device="cuda"); c = torch.empty(N * N, device="cuda")
run("matmul 8192", lambda: trn.check(L.trn_matmul_f32_dev(a.data_ptr(), N, N, b.data_ptr(), N, N, c.data_ptr(), st)))
x = torch.rand(1 << 28, device="cuda"); y = torch.empty_like(x)
run("gelu 2^28", lambda: trn.check(L.trn_gelu_f32_dev(x.data_ptr(), x.numel(), y.data_ptr(), st)))
