"""The rest of Vector's element-wise / statistics API on the CUDA path (through the C ABI) against the reference's own
known-answer tests and against the oracle on seeded data.

Contract (stated per op):
  * neg, signum, trunc, fract, hardswish, leaky_relu, clip, minimum, maximum, copysign, minmax_normalize: bit-exact
    against the oracle (single IEEE operations, or the reference's unfused operation order).
  * sinh, cosh, asin, acos, atan, asinh, acosh, atanh, elu, selu: <= 4 ulp against the f64 truth (CUDA's accurate
    functions are documented <= 2-4 ulp; glibc, which the reference calls through Rust std, <= 1-2 ulp); the oracle is
    held to the same bound.  pow: <= 8 ulp.  mish: <= 6 ulp + 2^-24 |x| (three chained transcendentals).
  * sum_of_squares, covariance, correlation, zscore, layer_norm_simple: the reduction contract of test_parity_gpu.py
    (1e-5 of the sum of magnitudes) carried through the reference's composition."""
import numpy as np
import pytest

import vector_api_kats

pytestmark = pytest.mark.gpu
f32 = np.float32


def ulp(x):
    return np.spacing(np.abs(np.asarray(x)).astype(f32)).astype(np.float64)


def _trn_call(trn):
    V = trn.Vector

    def call(op, *vecs, p=()):
        v = V.from_slice(np.asarray(vecs[0], f32))
        if op in ("minimum", "maximum", "copysign"):
            return getattr(v, op)(V.from_slice(np.asarray(vecs[1], f32))).as_slice()
        if op in ("covariance", "correlation"):
            return getattr(v, op)(V.from_slice(np.asarray(vecs[1], f32)))
        r = getattr(v, op)(*p)
        return r.as_slice() if isinstance(r, V) else r
    return call


def test_reference_kats_on_cuda(trn):
    vector_api_kats.run(_trn_call(trn))


def test_error_messages(trn):
    V = trn.Vector
    with pytest.raises(trn.TruenoError) as e:
        V.from_slice([1, 2, 3]).leaky_relu(1.5)
    assert e.value == trn.TruenoError.InvalidInput("negative_slope must be in [0.0, 1.0), got 1.5")   # src/vector.rs:1986-1990
    with pytest.raises(trn.TruenoError) as e:
        V.from_slice([1, 2, 3]).elu(-1.0)
    assert e.value == trn.TruenoError.InvalidInput("alpha must be > 0, got -1")                      # :2091-2095
    with pytest.raises(trn.TruenoError) as e:
        V.from_slice([1, 2, 3]).clip(10.0, 5.0)
    assert e.value == trn.TruenoError.InvalidInput("min_val (10) must be <= max_val (5)")           # :1449-1454
    with pytest.raises(trn.TruenoError) as e:
        V.from_slice([1, 2]).maximum(V.from_slice([1, 2, 3]))
    assert e.value == trn.TruenoError.SizeMismatch(2, 3)


SPECIALS = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 1.0, -1.0, 3.0, -3.0, 2.9999998, -2.9999998, 20.0, -20.0, 20.000002,
                     -20.000002, 1e-38, -1e-38, 1e-45, 3.4e38, -3.4e38, 0.5, -0.5, 88.0, -88.0, 100.0, -100.0], f32)


@pytest.mark.parametrize("n", [1, 7, 4099, 1 << 20])
def test_bit_exact_maps(trn, oracle, n):
    rng = np.random.default_rng(n)
    x = np.concatenate([SPECIALS, (rng.standard_normal(n) * 4).astype(f32)])
    y = np.concatenate([SPECIALS[::-1], (rng.standard_normal(n) * 4).astype(f32)])
    V = trn.Vector
    vx, vy = V.from_slice(x), V.from_slice(y)

    def same(a, b):   # bit patterns, NaN == NaN of any payload
        a, b = np.asarray(a, f32), np.asarray(b, f32)
        return np.array_equal(a.view(np.uint32)[~np.isnan(a)], b.view(np.uint32)[~np.isnan(b)]) and \
            np.array_equal(np.isnan(a), np.isnan(b))
    for op in ("neg", "signum", "trunc", "fract", "hardswish"):
        assert same(getattr(vx, op)().as_slice(), oracle.vector_map(op, x)), op
    assert same(vx.leaky_relu(0.01).as_slice(), oracle.vector_map("leaky_relu", x, p0=0.01))
    assert same(vx.leaky_relu(0.0).as_slice(), oracle.vector_map("leaky_relu", x, p0=0.0))
    assert same(vx.clip(-1.5, 2.25).as_slice(), oracle.vector_map("clip", x, p0=-1.5, p1=2.25))
    assert same(vx.copysign(vy).as_slice(), oracle.vector_map("copysign", x, y))
    # f32::min / max leave the sign of a +-0 tie unspecified: compare as values there
    for op in ("minimum", "maximum"):
        got, want = getattr(vx, op)(vy).as_slice(), oracle.vector_map(op, x, y)
        assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)]), op


@pytest.mark.parametrize("n", [5, 4099, 1 << 20])
def test_transcendental_maps_vs_truth(trn, oracle, n):
    rng = np.random.default_rng(n + 1)
    x = (rng.standard_normal(n) * 3).astype(f32)
    u = rng.uniform(-0.999, 0.999, n).astype(f32)
    V = trn.Vector
    with np.errstate(all="ignore"):
        cases = [("sinh", x, np.sinh, 4), ("cosh", x, np.cosh, 4), ("asin", u, np.arcsin, 4), ("acos", u, np.arccos, 4),
                 ("atan", x, np.arctan, 4), ("asinh", x, np.arcsinh, 4), ("acosh", (np.abs(x) + f32(1)).astype(f32), np.arccosh, 4),
                 ("atanh", u, np.arctanh, 4)]
        for op, arg, fn, k in cases:
            truth = fn(arg.astype(np.float64))
            got = getattr(V.from_slice(arg), op)().as_slice().astype(np.float64)
            assert np.all(np.abs(got - truth) <= k * ulp(truth) + 1e-45), op
            assert np.all(np.abs(oracle.vector_map(op, arg).astype(np.float64) - truth) <= k * ulp(truth) + 1e-45), op
        x64 = x.astype(np.float64)
        em1 = np.exp(x.astype(f32)).astype(np.float64)    # the reference rounds exp(x) to f32 before the subtraction
        # elu / selu: alpha * (exp(x) - 1): error of expf (<= 2 ulp of exp(x) <= 1) dominates for x < 0
        got = V.from_slice(x).elu(1.5).as_slice().astype(np.float64)
        want = oracle.vector_map("elu", x, p0=1.5).astype(np.float64)
        assert np.all(np.abs(got - want) <= 1.5 * 4 * 2.0 ** -24 + 2 * ulp(want))
        assert np.array_equal(got[x > 0], x64[x > 0])
        got = V.from_slice(x).selu().as_slice().astype(np.float64)
        want = oracle.vector_map("selu", x).astype(np.float64)
        assert np.all(np.abs(got - want) <= 1.76 * 4 * 2.0 ** -24 + 2 * ulp(want))
        # mish vs f64 truth of the reference's formula
        sp = np.log1p(np.exp(x64))
        truth = np.where(x64 < -20, 0.0, np.where(x64 > 20, x64, x64 * np.tanh(sp)))
        got = V.from_slice(x).mish().as_slice().astype(np.float64)
        assert np.all(np.abs(got - truth) <= 6 * ulp(truth) + 2.0 ** -24 * np.abs(x64))
        assert np.all(np.abs(oracle.vector_map("mish", x).astype(np.float64) - truth) <= 6 * ulp(truth) + 2.0 ** -24 * np.abs(x64))
        # pow
        base = (np.abs(x) + f32(0.01)).astype(f32)
        for e in (2.0, 0.5, -1.5, 3.0):
            truth = np.power(base.astype(np.float64), e)
            got = V.from_slice(base).pow(e).as_slice().astype(np.float64)
            assert np.all(np.abs(got - truth) <= 8 * ulp(truth)), e
    del em1


@pytest.mark.parametrize("n", [10, 1000, 100_003, 1 << 21])
def test_statistics_vs_oracle_and_truth(trn, oracle, n):
    rng = np.random.default_rng(n + 2)
    x = (rng.standard_normal(n) * 2 + 0.5).astype(f32)
    y = (0.3 * x + rng.standard_normal(n)).astype(f32)
    V = trn.Vector
    vx, vy = V.from_slice(x), V.from_slice(y)
    xd, yd = x.astype(np.float64), y.astype(np.float64)
    # sum_of_squares: the dot contract
    assert abs(float(vx.sum_of_squares()) - float(np.dot(xd, xd))) <= 1e-5 * float(np.dot(xd, xd))
    # covariance / correlation: E[xy] - mean_x mean_y computed in f32 from f32 moments: error ~ eps * (E|xy| + |mx my|)
    scale = float(np.mean(np.abs(xd * yd)) + abs(xd.mean() * yd.mean()))
    cov = float(np.mean(xd * yd) - xd.mean() * yd.mean())
    assert abs(float(vx.covariance(vy)) - cov) <= 2e-5 * scale
    assert abs(float(oracle.covariance(x, y)) - cov) <= 1e-4 * scale          # the AVX2 reference on the same scale
    corr = cov / (xd.std() * yd.std())
    assert abs(float(vx.correlation(vy)) - corr) <= 1e-4
    assert abs(float(vx.correlation(vx)) - 1.0) <= 1e-4                       # proptest: self-correlation is one
    # zscore: (x - mean) / std
    z = vx.zscore().as_slice().astype(np.float64)
    truth = (xd - xd.mean()) / xd.std()
    assert np.all(np.abs(z - truth) <= 2e-5 * (np.abs(truth) + 1))
    assert abs(z.mean()) < 1e-4 and abs(z.std() - 1) < 1e-4                   # proptests src/vector.rs:13176-13230
    # minmax_normalize: bit-exact (exact min / max, two rounded operations)
    assert np.array_equal(vx.minmax_normalize().as_slice(), oracle.minmax_normalize(x))
    # layer_norm_simple
    ln = vx.layer_norm_simple(1e-5).as_slice().astype(np.float64)
    truth = (xd - xd.mean()) / np.sqrt(xd.var() + 1e-5)
    assert np.all(np.abs(ln - truth) <= 2e-5 * (np.abs(truth) + 1))
    if n <= 100_003:
        assert np.all(np.abs(ln - oracle.layer_norm_simple(x, 1e-5)) <= 4e-5 * (np.abs(truth) + 1))


def test_layer_norm_simple_rows_dev_and_affine_dev(trn):
    torch = pytest.importorskip("torch")
    trn.check(trn.lib.trn_cuda_init(0))
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream or 1
    for rows, cols in [(64, 768), (33, 4096), (7, 12288), (5, 16384), (3, 20000), (4, 1001)]:
        x = torch.randn(rows, cols, device=dev) * 3 + 1
        y = torch.empty_like(x)
        trn.check(trn.lib.trn_layer_norm_simple_rows_f32_dev(x.data_ptr(), 1e-5, y.data_ptr(), rows, cols, st))
        torch.cuda.synchronize()
        xd = x.double()
        truth = (xd - xd.mean(1, keepdim=True)) / torch.sqrt(xd.var(1, unbiased=False, keepdim=True) + 1e-5)
        assert float(((y.double() - truth).abs() / (truth.abs() + 1)).max()) <= 2e-5
    x = torch.randn(1 << 20, device=dev)
    y = torch.empty_like(x)
    trn.check(trn.lib.trn_affine_f32_dev(x.data_ptr(), x.numel(), 0.25, 3.0, y.data_ptr(), st))
    torch.cuda.synchronize()
    assert torch.equal(y, (x - 0.25) * 3.0)


def test_embedding_lookup_reference_kats(trn):
    """src/matrix.rs:3727-3848"""
    M = trn.Matrix
    r = M.from_vec(4, 3, np.arange(1, 13)).embedding_lookup([1, 3, 0])
    assert r.shape() == (3, 3) and np.array_equal(r.as_slice(), np.array([4, 5, 6, 10, 11, 12, 1, 2, 3], f32))
    r = M.from_vec(3, 2, [1, 2, 3, 4, 5, 6]).embedding_lookup([1])
    assert r.shape() == (1, 2) and r.get(0, 0) == 3.0 and r.get(0, 1) == 4.0
    r = M.from_vec(2, 3, [1, 2, 3, 4, 5, 6]).embedding_lookup([0, 0, 1, 0])
    assert r.shape() == (4, 3) and r.get(0, 0) == r.get(1, 0) == r.get(3, 0)
    assert M.from_vec(3, 2, [1, 2, 3, 4, 5, 6]).embedding_lookup([]).shape() == (0, 2)
    with pytest.raises(trn.TruenoError) as e:
        M.from_vec(3, 2, [1, 2, 3, 4, 5, 6]).embedding_lookup([0, 5, 1])
    assert e.value == trn.TruenoError.InvalidInput("Index 5 at position 1 is out of bounds for embedding table with 3 rows")
    emb, uniq = M.from_vec(4, 2, [1, 2, 3, 4, 5, 6, 7, 8]).embedding_lookup_sparse([1, 3, 1, 0, 3])
    assert emb.shape() == (5, 2) and uniq == [0, 1, 3]
    r = M.from_vec(1000, 256, np.arange(1000 * 256)).embedding_lookup([0, 500, 999, 42, 100])
    assert r.shape() == (5, 256) and r.get(0, 0) == 0 and r.get(1, 0) == 500 * 256 and r.get(2, 0) == 999 * 256


@pytest.mark.parametrize("rows,cols,n", [(7, 1, 33), (50, 3, 1000), (1000, 64, 5000), (4000, 768, 3000), (300, 1001, 700),
                                         (128, 4096, 500), (64, 20000, 40), (9, 100_003, 12)])
def test_embedding_lookup_bit_exact_vs_oracle(trn, oracle, rows, cols, n):
    rng = np.random.default_rng(rows + cols + n)
    table = rng.standard_normal(rows * cols).astype(f32)
    table[:3] = [np.nan, np.inf, -0.0][: min(3, table.size)]
    idx = rng.integers(0, rows, n)
    got = trn.Matrix.from_vec(rows, cols, table).embedding_lookup(idx).as_slice()
    want = oracle.embedding_lookup(table, rows, cols, idx).reshape(-1)
    assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_embedding_lookup_dev_resident_and_out_of_range(trn):
    torch = pytest.importorskip("torch")
    trn.check(trn.lib.trn_cuda_init(0))
    dev = torch.device("cuda", 0)
    st = torch.cuda.current_stream().cuda_stream or 1
    for rows, cols in [(32000, 4096), (50257, 768), (1000, 1001)]:
        table = torch.randn(rows, cols, device=dev)
        idx = torch.randint(0, rows, (8192,), device=dev, dtype=torch.int64)
        idx[5] = rows + 7          # resident twin: an out-of-range index yields a zero row
        out = torch.full((idx.numel(), cols), 9.0, device=dev)
        trn.check(trn.lib.trn_embedding_lookup_f32_dev(table.data_ptr(), rows, cols, idx.data_ptr(), idx.numel(), out.data_ptr(), st))
        torch.cuda.synchronize()
        ok = torch.ones(idx.numel(), dtype=torch.bool, device=dev)
        ok[5] = False
        assert torch.equal(out[ok], table[idx[ok]]) and bool((out[5] == 0).all())


def test_cuda_path_against_committed_vector_api_golden(trn):
    """CUDA path vs the frozen golden file tests/golden/vector_api_fixtures.npz (oracle outputs on the xorshift fixtures) —
    no oracle call here, so an oracle change cannot mask a kernel change."""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vector_api_fixtures.npz"))
    x, y, u = g["x"], g["y"], g["u"]
    V = trn.Vector
    vx, vy, vu = V.from_slice(x), V.from_slice(y), V.from_slice(u)
    for op in ("neg", "signum", "trunc", "fract", "hardswish"):                                  # bit-exact
        assert np.array_equal(getattr(vx, op)().as_slice(), g[op]), op
    assert np.array_equal(vx.leaky_relu(0.01).as_slice(), g["leaky_relu_0.01"])
    assert np.array_equal(vx.clip(-0.5, 0.75).as_slice(), g["clip_-0.5_0.75"])
    for op in ("minimum", "maximum", "copysign"):
        assert np.array_equal(getattr(vx, op)(vy).as_slice(), g[op]), op
    assert np.array_equal(vx.minmax_normalize().as_slice(), g["minmax_normalize"])
    for op, k in (("sinh", 4), ("cosh", 4), ("atan", 4), ("asinh", 6)):                         # CUDA vs glibc: <= k ulp apart
        want = g[op].astype(np.float64)
        assert np.all(np.abs(getattr(vx, op)().as_slice() - want) <= k * ulp(want)), op
    for op, k in (("asin", 4), ("acos", 4), ("atanh", 6)):
        want = g[op].astype(np.float64)
        assert np.all(np.abs(getattr(vu, op)().as_slice() - want) <= k * ulp(want)), op
    want = g["acosh"].astype(np.float64)
    assert np.all(np.abs(V.from_slice((np.abs(x) + f32(1)).astype(f32)).acosh().as_slice() - want) <= 6 * ulp(want) + 1e-7)
    want = g["pow_2"].astype(np.float64)
    assert np.all(np.abs(vx.pow(2.0).as_slice() - want) <= 4 * ulp(want))
    for op, got, amp in (("mish", vx.mish().as_slice(), 1.0), ("selu", vx.selu().as_slice(), 1.76), ("elu_1.5", vx.elu(1.5).as_slice(), 1.5)):
        want = g[op].astype(np.float64)
        assert np.all(np.abs(got - want) <= 8 * ulp(want) + amp * 8 * 2.0 ** -24), op
    want = g["zscore"].astype(np.float64)
    assert np.all(np.abs(vx.zscore().as_slice() - want) <= 2e-5 * (np.abs(want) + 1))
    want = g["layer_norm_simple_1e-5"].astype(np.float64)
    assert np.all(np.abs(vx.layer_norm_simple(1e-5).as_slice() - want) <= 2e-5 * (np.abs(want) + 1))
    ss, cov, corr = (float(v) for v in g["stats"])
    assert abs(float(vx.sum_of_squares()) - ss) <= 1e-5 * ss
    scale = float(np.mean(np.abs(x.astype(np.float64) * y)) + abs(x.mean() * y.mean()))
    assert abs(float(vx.covariance(vy)) - cov) <= 4e-5 * scale and abs(float(vx.correlation(vy)) - corr) <= 1e-4
    got = trn.Matrix.from_vec(64, 32, x).embedding_lookup(g["embedding_idx"]).as_slice().reshape(-1, 32)
    assert np.array_equal(got, g["embedding_64x32"])
