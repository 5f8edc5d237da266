"""Small invocations of the A-stationary fused-split GEMM (rotated tile order, ragged segments) and of the rotated matvec for
compute-sanitizer memcheck / synccheck."""
import sys
sys.path.insert(0, ".")
import numpy as np
import trueno_b200 as trn
trn.check(trn.lib.trn_cuda_init(0))
trn.set_gemm_engine(trn.ENGINE_TC_3XTF32)
rng = np.random.default_rng(3)
f32 = np.float32
for batch, m, k, n in [(2, 520, 128, 1100), (5, 300, 100, 772), (1, 1024, 128, 2048), (37, 256, 32, 512), (1, 256, 4, 516)]:
    A = rng.uniform(-1, 1, (batch, m, k)).astype(f32); B = rng.uniform(-1, 1, (batch, k, n)).astype(f32)
    C = trn.Matrix.batched_matmul(A.ravel(), B.ravel(), batch, m, k, n).reshape(batch, m, n)
    truth = A.astype(np.float64) @ B.astype(np.float64)
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)
    assert np.all(np.abs(C - truth) <= 1e-5 * scale), (batch, m, k, n)
trn.set_gemm_engine(0)
for rows, cols in [(300, 16384), (64, 32768 + 4096), (20, 65536)]:
    A = rng.uniform(-1, 1, (rows, cols)).astype(f32); v = rng.uniform(-1, 1, cols).astype(f32)
    y = trn.Matrix.from_vec(rows, cols, A.ravel()).matvec(trn.Vector.from_slice(v)).as_slice()
    ty = A.astype(np.float64) @ v.astype(np.float64); sy = np.abs(A).astype(np.float64) @ np.abs(v).astype(np.float64)
    assert np.all(np.abs(y - ty) <= 1e-5 * sy), (rows, cols)
print("sanitize gemm/matvec workload ok")
