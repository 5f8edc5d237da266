"""PCIe floor for the e2e matmul step: pinned H2D 512 MiB, D2H 256 MiB, alone and concurrently."""
import time, torch
h_in = torch.empty(128 << 20, dtype=torch.float32).pin_memory()
h_out = torch.empty(64 << 20, dtype=torch.float32).pin_memory()
d_in = torch.empty(128 << 20, dtype=torch.float32, device="cuda")
d_out = torch.empty(64 << 20, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(h2d, d2h):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5):
        if h2d:
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
        if d2h:
            with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / 5 * 1e3
for _ in range(2): run(True, True)
a, b, c = run(True, False), run(False, True), run(True, True)
print(f"H2D 512 MiB alone {a:.2f} ms ({536.9/a:.1f} GB/s)   D2H 256 MiB alone {b:.2f} ms ({268.4/b:.1f} GB/s)   both {c:.2f} ms")
