"""The reference's proptest suite (100 cases each) restated with hypothesis against the CUDA path:
src/vector.rs:9030-14280 (vector properties), src/matrix.rs:3251-3507 (matrix properties).  Same generators
(values in -1000..1000 / -100..100, lengths 1..100, small matrix dims), same tolerances."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st
from hypothesis.extra import numpy as hnp

pytestmark = pytest.mark.gpu
f32 = np.float32
CASES = settings(max_examples=100, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])


def vec(lo=-1000.0, hi=1000.0, min_len=1, max_len=100):
    return hnp.arrays(f32, st.integers(min_len, max_len), elements=st.floats(lo, hi, width=32, allow_nan=False))


@CASES
@given(data=st.data())
def test_dot_commutative_and_norm_is_sqrt_dot(trn, data):       # src/vector.rs:9100-9180, :11830-11870
    n = data.draw(st.integers(1, 100))
    a = data.draw(hnp.arrays(f32, n, elements=st.floats(-100, 100, width=32)))
    b = data.draw(hnp.arrays(f32, n, elements=st.floats(-100, 100, width=32)))
    va, vb = trn.Vector(a), trn.Vector(b)
    assert va.dot(vb) == vb.dot(va)
    assert abs(float(va.norm_l2()) - float(np.sqrt(va.dot(va)))) <= 1e-3 * max(1.0, float(va.norm_l2()))


@CASES
@given(a=vec())
def test_sum_max_min_match_manual(trn, a):                        # src/vector.rs:9200-9330
    v = trn.Vector(a)
    manual = float(np.sum(a.astype(np.float64)))
    assert abs(float(v.sum()) - manual) <= 1e-3 * max(1.0, float(np.abs(a).sum()))
    assert float(v.max()) == float(a.max()) and float(v.min()) == float(a.min())


@CASES
@given(a=hnp.arrays(f32, st.integers(1, 100), elements=st.integers(-1000, 1000).map(float)))
def test_argmax_argmin_first_occurrence(trn, a):                  # src/vector.rs:9434-9477
    v = trn.Vector(a)
    assert v.argmax() == int(np.argmax(a)) and v.argmin() == int(np.argmin(a))   # numpy also returns the first


@CASES
@given(a=vec(-50.0, 50.0), shift=st.floats(-10, 10, width=32))
def test_softmax_sums_to_one_and_translation_invariant(trn, a, shift):   # src/vector.rs:13461-13530
    s = trn.Vector(a).softmax().as_slice()
    assert abs(float(s.sum(dtype=np.float64)) - 1.0) < 1e-5
    assert ((s >= 0) & (s <= 1)).all()
    s2 = trn.Vector((a + f32(shift)).astype(f32)).softmax().as_slice()
    assert np.max(np.abs(s - s2)) < 1e-4


@CASES
@given(a=vec(-100.0, 100.0))
def test_sigmoid_bounded_and_monotone_relu_gelu(trn, a):           # src/vector.rs:13668-13714, :13240-13300
    s = trn.Vector(a).sigmoid().as_slice()
    assert ((s >= 0) & (s <= 1)).all()
    order = np.argsort(a, kind="stable")
    assert np.all(np.diff(s[order]) >= 0)
    r = trn.Vector(a).relu().as_slice()
    assert np.array_equal(r, np.maximum(a, 0).astype(f32) + f32(0))
    g = trn.Vector(a).gelu().as_slice()
    big = a > 10
    assert np.all(np.abs(g[big] - a[big]) <= 1e-3 * np.abs(a[big]))          # linear for large x (src/vector.rs:8397)


def small_matrix(rows, cols, lo=-10.0, hi=10.0):
    return hnp.arrays(f32, (rows, cols), elements=st.floats(lo, hi, width=32))


@CASES
@given(data=st.data())
def test_matmul_identity_associativity_transpose(trn, data):      # src/matrix.rs:3267-3420
    m, k, n, p = (data.draw(st.integers(1, 8)) for _ in range(4))
    A, B, C = data.draw(small_matrix(m, k)), data.draw(small_matrix(k, n)), data.draw(small_matrix(n, p))
    M = trn.Matrix
    mA, mB, mC = M.from_vec(m, k, A), M.from_vec(k, n, B), M.from_vec(n, p, C)
    assert np.max(np.abs(mA.matmul(M.identity(k)).to_numpy() - A)) <= 1e-5                      # A * I = A
    lhs = mA.matmul(mB).matmul(mC).to_numpy()
    rhs = mA.matmul(mB.matmul(mC)).to_numpy()
    scale = np.abs(A).astype(np.float64) @ np.abs(B) @ np.abs(C) + 1e-6
    assert np.all(np.abs(lhs - rhs) <= 0.05 * scale)                                            # (AB)C = A(BC), 5 %
    t1 = mA.matmul(mB).transpose().to_numpy()
    t2 = mB.transpose().matmul(mA.transpose()).to_numpy()
    assert np.all(np.abs(t1 - t2) <= 1e-3 * np.maximum(np.abs(t1), 1))                          # (AB)^T = B^T A^T


@CASES
@given(data=st.data())
def test_matvec_and_vecmat_associativity(trn, data):               # src/matrix.rs:3430-3507
    m, k, n = (data.draw(st.integers(1, 8)) for _ in range(3))
    A, B = data.draw(small_matrix(m, k)), data.draw(small_matrix(k, n))
    v = data.draw(hnp.arrays(f32, n, elements=st.floats(-10, 10, width=32)))
    w = data.draw(hnp.arrays(f32, m, elements=st.floats(-10, 10, width=32)))
    M, V = trn.Matrix, trn.Vector
    mA, mB = M.from_vec(m, k, A), M.from_vec(k, n, B)
    scale = np.abs(A).astype(np.float64) @ np.abs(B) @ np.abs(v) + 1e-6
    lhs = mA.matmul(mB).matvec(V(v)).as_slice()
    rhs = mA.matvec(mB.matvec(V(v))).as_slice()
    assert np.all(np.abs(lhs - rhs) <= 2e-2 * scale)                                            # (AB)v = A(Bv)
    scale2 = np.abs(w).astype(np.float64) @ np.abs(A) @ np.abs(B) + 1e-6
    l2 = M.vecmat(V(w), mA.matmul(mB)).as_slice()
    r2 = M.vecmat(M.vecmat(V(w), mA), mB).as_slice()
    assert np.all(np.abs(l2 - r2) <= 2e-2 * scale2)                                             # w(AB) = (wA)B


# ---- widened rows: SymmetricEigen (the reference's proptests, src/eigen.rs:805-870) and the fused attention -----------
@settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(data=st.data())
def test_eigen_properties(trn, data):
    """prop_eigenvalues_descending (n in 2..6), prop_eigenvector_count_matches_dimension (1..8),
    prop_reconstruction_accuracy — here for n in 1..24 and arbitrary symmetric matrices."""
    n = data.draw(st.integers(1, 24))
    a = data.draw(hnp.arrays(f32, (n, n), elements=st.floats(-10, 10, width=32)))
    m = ((a + a.T) / 2).astype(f32)
    eig = trn.SymmetricEigen.new(trn.Matrix.from_vec(n, n, m.ravel()))
    vals = eig.eigenvalues().astype(np.float64)
    vecs = eig.eigenvectors().as_slice().reshape(n, n).astype(np.float64)
    frob = max(1.0, float(np.linalg.norm(m.astype(np.float64))))
    assert len(eig) == n and vecs.shape == (n, n)
    assert np.all(np.diff(vals) <= 0)
    assert np.max(np.abs(vecs @ np.diag(vals) @ vecs.T - m)) <= 1e-5 * frob * max(1.0, n ** 0.5)
    assert np.max(np.abs(vals - np.linalg.eigvalsh(np.triu(m).astype(np.float64) + np.triu(m, 1).T.astype(np.float64))[::-1])) <= 4e-6 * frob


@settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])
@given(data=st.data())
def test_attention_properties(trn, data):
    """Rows of the output are convex combinations of the value rows (inside their min/max per column), and the result
    matches the f64 truth within the stated contract — random ragged shapes, both masks, both kernel forms' ranges."""
    heads = data.draw(st.integers(1, 3))
    seq = data.draw(st.integers(1, 300))
    d = data.draw(st.sampled_from([1, 3, 8, 16, 24, 40, 64, 72, 128, 136]))
    causal = data.draw(st.booleans())
    seed = data.draw(st.integers(0, 2 ** 31))
    rng = np.random.default_rng(seed)
    q, k, v = (rng.standard_normal(heads * seq * d).astype(f32) * f32(1.5) for _ in range(3))
    scale = f32(1.0) / np.sqrt(f32(d))
    got = trn.attention(q, k, v, heads, seq, d, causal=causal).reshape(heads, seq, d)
    q64, k64, v64 = (x.reshape(heads, seq, d).astype(np.float64) for x in (q, k, v))
    s = np.einsum("hid,hjd->hij", q64, k64) * float(scale)
    if causal:
        s = np.where(np.arange(seq)[None, None, :] > np.arange(seq)[None, :, None], -np.inf, s)
    p = np.exp(s - s.max(axis=-1, keepdims=True))
    p /= p.sum(axis=-1, keepdims=True)
    want = np.einsum("hij,hjd->hid", p, v64)
    bound = np.einsum("hij,hjd->hid", p, np.abs(v64))
    kappa = float(scale) * np.einsum("hid,hjd->hij", np.abs(q64), np.abs(k64)).max()
    assert np.all(np.abs(got - want) <= (1e-5 + 2e-5 * kappa) * bound + 1e-30)
    lo, hi = v64.min(axis=1, keepdims=True), v64.max(axis=1, keepdims=True)
    span = (hi - lo) + np.abs(hi) + np.abs(lo)
    assert np.all(got >= lo - 1e-5 * span) and np.all(got <= hi + 1e-5 * span)


# ---- statistics and activations of the widened Vector API (src/vector.rs:12763-13330 and the per-op proptests) ----
@CASES
@given(a=vec(-100.0, 100.0), k=st.floats(-5, 5, width=32))
def test_sum_of_squares_properties(trn, a, k):                      # src/vector.rs:12763-12820
    v = trn.Vector(a)
    ss = float(v.sum_of_squares())
    assert ss >= 0.0
    assert abs(ss - float(v.dot(v))) < 1e-4 * max(1.0, ss)
    small = (a / f32(10)).astype(f32)
    vs = trn.Vector(small)
    want = float(k) * float(k) * float(vs.sum_of_squares())
    assert abs(float(vs.scale(float(k)).sum_of_squares()) - want) < 1e-3 * max(abs(want), 1.0)


@CASES
@given(data=st.data())
def test_covariance_and_correlation_properties(trn, data):          # src/vector.rs:13020-13170
    n = data.draw(st.integers(2, 100))
    a = data.draw(hnp.arrays(f32, n, elements=st.floats(-50, 50, width=32)))
    b = data.draw(hnp.arrays(f32, n, elements=st.floats(-50, 50, width=32)))
    va, vb = trn.Vector(a), trn.Vector(b)
    var = float(va.variance())
    assert abs(float(va.covariance(va)) - var) < 1e-3 * max(abs(var), 1e-5) + 1e-3        # Cov(X, X) = Var(X)
    cab, cba = float(va.covariance(vb)), float(vb.covariance(va))
    assert abs(cab - cba) <= 1e-4 * max(abs(cab), 1e-5) + 1e-4                             # symmetry
    # E[x^2] - mean^2 in f32 (the reference's formula, src/vector.rs:983-995) cancels for nearly constant vectors:
    # correlation is only meaningful (in the reference too) when the spread is not tiny against the magnitude
    if min(a.std() / max(1.0, np.abs(a).max()), b.std() / max(1.0, np.abs(b).max())) < 0.05:
        return
    try:
        r = float(va.correlation(vb))
    except trn.TruenoError as e:                                                           # constant vector
        assert e.variant == "DivisionByZero"
        return
    assert -1.0 <= r <= 1.0                                                                # bounded
    assert abs(r - float(vb.correlation(va))) < 1e-3                                       # symmetric


@CASES
@given(a=vec(-100.0, 100.0, min_len=2))
def test_zscore_and_minmax_properties(trn, a):                      # src/vector.rs:13176-13330
    v = trn.Vector(a)
    if a.std() / max(1.0, np.abs(a).max()) < 0.05:      # see the note on E[x^2] - mean^2 above
        return
    z = v.zscore().as_slice().astype(np.float64)
    assert abs(z.mean()) < 1e-3 and abs(z.std() - 1.0) < 2e-2
    m = v.minmax_normalize().as_slice()
    assert float(m.min()) == 0.0 and abs(float(m.max()) - 1.0) < 1e-5 and ((m >= 0) & (m <= 1.0 + 1e-6)).all()
    assert np.all(np.diff(m[np.argsort(a, kind="stable")]) >= 0)                           # order preserving


@CASES
@given(a=vec(-100.0, 100.0), slope=st.floats(0.0, 0.984375, width=32), alpha=st.floats(0.0625, 5.0, width=32))
def test_activation_properties(trn, a, slope, alpha):               # src/vector.rs: leaky_relu / elu / hardswish / mish / selu proptests
    v = trn.Vector(a)
    pos = a > 0
    lr = v.leaky_relu(float(slope)).as_slice()
    assert np.array_equal(lr[pos], a[pos]) and np.array_equal(lr[~pos], (f32(slope) * a[~pos]).astype(f32))
    el = v.elu(float(alpha)).as_slice()
    assert np.array_equal(el[pos], a[pos]) and (el[~pos] >= -f32(alpha) * (1 + 1e-6)).all() and (el[~pos] <= 0).all()
    hs = v.hardswish().as_slice()
    assert np.array_equal(hs[a >= 3], a[a >= 3]) and (hs[a <= -3] == 0).all() and (hs >= -0.375 - 1e-6).all()
    mi = v.mish().as_slice()
    assert (mi >= -0.31).all() and np.all(np.abs(mi[a > 20] - a[a > 20]) == 0) and (mi[a < -20] == 0).all()
    se = v.selu().as_slice()
    assert (se[pos] > 0).all() and (se[~pos] <= 0).all() and (se >= -1.7581 - 1e-4).all()
    # clip is idempotent and bounded; minimum / maximum bracket their operands
    c = v.clip(-10.0, 10.0)
    assert np.array_equal(c.clip(-10.0, 10.0).as_slice(), c.as_slice()) and np.abs(c.as_slice()).max() <= 10.0
    w = trn.Vector(a[::-1].copy())
    lo, hi = v.minimum(w).as_slice(), v.maximum(w).as_slice()
    assert (lo <= hi).all() and np.array_equal(lo + hi, a + a[::-1])
    assert np.array_equal(v.neg().neg().as_slice(), a) and np.array_equal(np.abs(v.signum().as_slice()), np.ones_like(a))
    assert np.array_equal(v.trunc().as_slice() + v.fract().as_slice(), a) or np.allclose(v.trunc().as_slice() + v.fract().as_slice(), a, rtol=0, atol=1e-4)


@CASES
@given(data=st.data())
def test_embedding_lookup_properties(trn, data):                    # gather: rows come back verbatim, order and repeats preserved
    rows = data.draw(st.integers(1, 40))
    cols = data.draw(st.integers(1, 70))
    table = data.draw(hnp.arrays(f32, rows * cols, elements=st.floats(-1000, 1000, width=32)))
    idx = data.draw(st.lists(st.integers(0, rows - 1), min_size=0, max_size=60))
    m = trn.Matrix.from_vec(rows, cols, table)
    got = m.embedding_lookup(idx)
    assert got.shape() == (len(idx), cols)
    assert np.array_equal(got.as_slice().reshape(len(idx), cols), table.reshape(rows, cols)[np.asarray(idx, dtype=np.int64)])
    emb, uniq = m.embedding_lookup_sparse(idx)
    assert uniq == sorted(set(idx)) and np.array_equal(emb.as_slice(), got.as_slice())
