"""Pattern probes of the fused-split GEMM: which (m, n, k) of the operands lands where.  TRN_GEMM_FUSED=1."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import trueno_b200 as trn  # noqa: E402

f32 = np.float32
trn.check(trn.lib.trn_cuda_init(0))
trn.set_gemm_engine(trn.ENGINE_TC_3XTF32)


def mm(A, B):
    m, k = A.shape
    n = B.shape[1]
    return trn.Matrix.from_vec(m, k, A).matmul(trn.Matrix.from_vec(k, n, B)).to_numpy()


def show(name, C, want):
    bad = np.argwhere(C != want)
    print(f"{name}: {'OK' if len(bad) == 0 else f'{len(bad)} wrong of {C.size}'}")
    if len(bad):
        print("   first wrong at", bad[:6].tolist())
        i, j = bad[0]
        print("   got ", C[i, max(0, j - 2):j + 6].tolist())
        print("   want", want[i, max(0, j - 2):j + 6].tolist())
        print("   corner values got", C[0, 0], C[0, -1], C[-1, 0], C[-1, -1], "want", want[0, 0], want[0, -1], want[-1, 0], want[-1, -1])
        rows_bad = np.unique(bad[:, 0]); cols_bad = np.unique(bad[:, 1])
        print(f"   bad rows {rows_bad[:8].tolist()}..{rows_bad[-1]} ({len(rows_bad)}), bad cols {cols_bad[:8].tolist()}..{cols_bad[-1]} ({len(cols_bad)})")


m, k, n = 256, 16, 256
ones_a, ones_b = np.ones((m, k), f32), np.ones((k, n), f32)
show("ones", mm(ones_a, ones_b), np.full((m, n), k, f32))
Bn = np.tile(np.arange(n, dtype=f32), (k, 1))
show("B=n", mm(ones_a, Bn), (ones_a.astype(np.float64) @ Bn).astype(f32))
Am = np.tile(np.arange(m, dtype=f32)[:, None], (1, k))
show("A=m", mm(Am, ones_b), (Am.astype(np.float64) @ ones_b).astype(f32))
for k0 in (0, 1, 7, 8, 15):
    A = np.zeros((m, k), f32); A[:, k0] = 1
    Bk = np.tile(np.arange(k, dtype=f32)[:, None], (1, n)) + 1
    show(f"A=e{k0}, B=k+1", mm(A, Bk), np.full((m, n), k0 + 1, f32))
m, k, n = 512, 64, 512
rng = np.random.default_rng(0)
A = rng.integers(0, 8, (m, k)).astype(f32); B = rng.integers(0, 8, (k, n)).astype(f32)
show("ints 512x64x512", mm(A, B), (A.astype(np.float64) @ B).astype(f32))
A = rng.uniform(0, 1, (m, k)).astype(f32); B = rng.uniform(0, 1, (k, n)).astype(f32)
C = mm(A, B); T = A.astype(np.float64) @ B.astype(np.float64)
print("uniform 512x64x512 max rel err", float(np.max(np.abs(C - T) / T)))
