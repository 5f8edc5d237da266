"""Small invocations of the kernels added late in round 1, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import sys
sys.path.insert(0, ".")
import numpy as np
import trueno_b200 as trn
rng = np.random.default_rng(0)
f32 = np.float32
for heads, seq, d, causal in [(2, 300, 64, False), (1, 384, 128, True), (1, 100, 32, False), (1, 50, 200, True)]:
    q, k, v = (rng.standard_normal(heads * seq * d).astype(f32) for _ in range(3))
    out = trn.attention(q, k, v, heads, seq, d, causal=causal)
    assert np.all(np.isfinite(out))
for n in (5, 64, 130):
    m = rng.standard_normal((n, n)).astype(f32); m = ((m + m.T) / 2).astype(f32)
    e = trn.SymmetricEigen.new(trn.Matrix.from_vec(n, n, m.ravel()))
    assert np.all(np.diff(e.eigenvalues()) <= 0)
a = rng.standard_normal((300, 200)).astype(f32); b = rng.standard_normal((200, 260)).astype(f32)
trn.set_gemm_engine(trn.ENGINE_TC_3XTF32)
c = trn.Matrix.from_vec(300, 200, a.ravel()).matmul(trn.Matrix.from_vec(200, 260, b.ravel()))
trn.set_gemm_engine(trn.ENGINE_AUTO)
print("sanitize workload ok")
