"""Phase times of the pipelined host-slice matmul (TRN_PIPE_TRACE=1) at 8192^3, pinned buffers."""
import os, sys, time
os.environ["TRN_PIPE_TRACE"] = "1"
sys.path.insert(0, ".")
import numpy as np
import trueno_b200 as trn
L = trn.lib
trn.check(L.trn_cuda_init(0))
n = 8192
ha, hb, hc = trn.pinned_empty(n * n), trn.pinned_empty(n * n), trn.pinned_empty(n * n)
ha[:] = np.random.default_rng(1).random(n * n, dtype=np.float32); hb[:] = ha[::-1]
for i in range(4):
    t0 = time.perf_counter()
    trn.check(L.trn_matmul_f32(ha.ctypes.data, n, n, hb.ctypes.data, n, n, hc.ctypes.data))
    print(f"call {i}: {(time.perf_counter() - t0) * 1e3:.2f} ms host clock", flush=True)
