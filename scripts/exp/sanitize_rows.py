"""Small invocations of every row-kernel family and form (aligned / window / long / split / sharded slices), the new
maps, the statistics and the gather, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python scripts/exp/sanitize_rows.py
    compute-sanitizer --tool racecheck python scripts/exp/sanitize_rows.py
    compute-sanitizer --tool synccheck python scripts/exp/sanitize_rows.py
Host-slice calls allocate exactly rows * cols elements on the device, so a window kernel touching a byte outside its
rows is an out-of-bounds access memcheck reports."""
import sys
sys.path.insert(0, ".")
import numpy as np
import trueno_b200 as trn
rng = np.random.default_rng(0)
f32 = np.float32
for rows, cols in [(3, 1), (5, 7), (9, 77), (4, 1000), (7, 1001), (5, 1019), (3, 4099), (2, 8191), (3, 8192), (2, 16387), (2, 20000),
                   (3, 32001), (2, 32768), (20, 40000), (19, 50257), (20, 65537), (19, 131072), (2, 70001), (1, 300_003), (1, 1 << 20)]:
    x = (rng.standard_normal((rows, cols)) * 4).astype(f32)
    for log in (False, True):
        y = trn.softmax_rows(x, rows, cols, log=log)
        assert np.all(np.isfinite(y))
V = trn.Vector
x = rng.standard_normal(10_007).astype(f32)
y = rng.standard_normal(10_007).astype(f32)
v, w = V.from_slice(x), V.from_slice(y)
for op in ("neg", "signum", "trunc", "fract", "sinh", "cosh", "atan", "asinh", "hardswish", "mish", "selu"):
    getattr(v, op)()
v.leaky_relu(0.1); v.elu(1.0); v.pow(2.0); v.clip(-1.0, 1.0); v.minimum(w); v.maximum(w); v.copysign(w)
v.zscore(); v.minmax_normalize(); v.covariance(w); v.correlation(w); v.sum_of_squares(); v.layer_norm_simple(1e-5)
for n in (8192, 12288, 16384, 20000):
    z = rng.standard_normal(n).astype(f32)
    V.from_slice(z).layer_norm(V.from_slice(np.ones(n, f32)), V.from_slice(np.zeros(n, f32)), 1e-5)
for rows, cols, n in [(50, 3, 100), (100, 64, 333), (64, 768, 100), (9, 1001, 40), (16, 20000, 8)]:
    t = rng.standard_normal(rows * cols).astype(f32)
    trn.Matrix.from_vec(rows, cols, t).embedding_lookup(rng.integers(0, rows, n))
print("sanitize rows workload ok")
