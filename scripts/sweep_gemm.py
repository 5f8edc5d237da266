"""Engine crossover sweep: SIMT FFMA vs tcgen05 3xTF32 for square and skinny shapes (device-resident, CUDA events)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import trueno_b200 as trn

def timeit(fn, iters=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / iters

torch.cuda.set_device(0); trn.check(trn.lib.trn_cuda_init(0))
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); st = stream.cuda_stream
L = trn.lib
shapes = [(s, s, s) for s in (128, 192, 256, 384, 512, 768, 1024, 1536, 2048)] + [(128, 4096, 128), (256, 256, 4096), (4096, 64, 4096), (2048, 128, 2048), (160, 8192, 160), (8192, 256, 256)]
for m, k, n in shapes:
    a, b, c = torch.rand(m, k, device="cuda"), torch.rand(k, n, device="cuda"), torch.empty(m, n, device="cuda")
    res = {}
    for name, eng in (("simt", 1), ("tc3", 2), ("auto", 0)):
        trn.set_gemm_engine(eng)
        res[name] = timeit(lambda: trn.check(L.trn_matmul_f32_dev(a.data_ptr(), m, k, b.data_ptr(), k, n, c.data_ptr(), st)))
    trn.set_gemm_engine(0)
    fl = 2.0 * m * k * n
    print(f"{m:5d}x{k:5d}x{n:5d}  simt {res['simt']*1e3:8.1f} us ({fl/res['simt']/1e9:6.1f} TF)  tc3 {res['tc3']*1e3:8.1f} us ({fl/res['tc3']/1e9:6.1f} TF)  auto {res['auto']*1e3:8.1f} us  -> {'tc3' if res['tc3'] < res['simt'] else 'simt'} wins")
