"""SymmetricEigen on the CUDA path (csrc/eigen.cu: parallel Jacobi) against the reference's own tests
(src/eigen.rs:526-870), the oracle (the reference's cyclic Jacobi) and the f64 LAPACK spectrum.

The device applies the reference's rotations in a different ORDER (n/2 disjoint pairs at a time), so parity is stated
as tolerances, all relative to ||A||_F (both implementations stop when every off-diagonal is below 1e-7 * ||A||_F):
eigenvalues within 2e-6 * ||A||_F of the f64 truth and of the oracle's; V^T V = I within 1e-5 * sqrt(n); A V = V L within
4e-6 * ||A||_F per element; eigenvectors of well-separated eigenvalues equal to the oracle's up to sign within 1e-4."""
import numpy as np
import pytest

from eigen_kats import COV_MATRIX, EIGEN_KATS, ORTHO_MATRIX

pytestmark = pytest.mark.gpu
f32 = np.float32


def sym(n, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    m = rng.standard_normal((n, n)) * scale
    return ((m + m.T) / 2).astype(f32)


@pytest.mark.parametrize("kat", EIGEN_KATS, ids=[k[3] for k in EIGEN_KATS])
def test_eigen_reference_kats(trn, kat):
    rows, want, tol, _ = kat
    n = len(rows)
    m = trn.Matrix.from_vec(n, n, np.asarray(rows, f32).ravel())
    eig = trn.SymmetricEigen.new(m)
    vals = eig.eigenvalues()
    assert len(eig) == n and not eig.is_empty()
    assert np.all(np.diff(vals) <= 0)
    assert np.max(np.abs(vals - np.asarray(want, f32))) < tol
    for lam, vec in eig:                                             # src/eigen.rs:644-668
        av = m.matvec(vec).as_slice()
        assert np.max(np.abs(av - lam * vec.as_slice())) < 1e-4
    rec = eig.reconstruct()                                          # src/eigen.rs:620-640
    assert np.max(np.abs(rec.as_slice() - m.as_slice())) < 1e-4
    assert eig.eigenvector(0).len() == n and eig.eigenvector(10) is None   # src/eigen.rs:726-735


def test_eigen_orthogonal_and_covariance(trn):
    m = trn.Matrix.from_vec(3, 3, np.asarray(ORTHO_MATRIX, f32).ravel())
    v = trn.SymmetricEigen.new(m).eigenvectors()
    prod = v.transpose().matmul(v).as_slice().reshape(3, 3)          # src/eigen.rs:591-616
    assert np.max(np.abs(prod - np.eye(3))) < 1e-4
    vals = trn.SymmetricEigen.new(trn.Matrix.from_vec(2, 2, np.asarray(COV_MATRIX, f32).ravel())).eigenvalues()
    assert vals[0] > 5.0 and abs(vals[1]) < 0.1                      # src/eigen.rs:739-750


def test_eigen_errors(trn):
    with pytest.raises(trn.TruenoError) as e:                        # src/eigen.rs:672-682
        trn.SymmetricEigen.new(trn.Matrix.from_vec(2, 3, np.arange(6, dtype=f32)))
    assert e.value.variant == "InvalidInput"
    assert "Matrix must be square for eigendecomposition, got 2x3" in str(e.value)
    with pytest.raises(trn.TruenoError) as e:                        # src/eigen.rs:686-690
        trn.SymmetricEigen.new(trn.Matrix.zeros(0, 0))
    assert "Cannot compute eigendecomposition of empty matrix" in str(e.value)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 7, 8, 16, 33, 64, 100, 257])
def test_eigen_vs_oracle_and_lapack(trn, oracle, n):
    m = sym(n, seed=n, scale=1.0 + n % 3)
    eig = trn.SymmetricEigen.new(trn.Matrix.from_vec(n, n, m.ravel()))
    vals = eig.eigenvalues().astype(np.float64)
    vecs = eig.eigenvectors().as_slice().reshape(n, n).astype(np.float64)
    frob = np.linalg.norm(m.astype(np.float64))
    truth = np.linalg.eigvalsh(m.astype(np.float64))[::-1]
    assert np.all(np.diff(vals) <= 0)
    assert np.max(np.abs(vals - truth)) <= 2e-6 * frob
    assert np.max(np.abs(vecs.T @ vecs - np.eye(n))) <= 1e-5 * max(1.0, n ** 0.5)
    assert np.max(np.abs(m.astype(np.float64) @ vecs - vecs * vals[None, :])) <= 4e-6 * frob
    ovals, ovecs = oracle.symmetric_eigen(m.ravel(), n, n)
    assert np.max(np.abs(vals - ovals)) <= 2e-6 * frob
    gaps = np.minimum(np.abs(np.diff(truth, prepend=np.inf)), np.abs(np.diff(truth, append=-np.inf)))
    for i in np.nonzero(gaps > 0.05 * frob / n ** 0.5)[0]:          # well-separated: same vector up to sign
        d = min(np.max(np.abs(vecs[:, i] - ovecs[:, i])), np.max(np.abs(vecs[:, i] + ovecs[:, i])))
        assert d <= 1e-4, (i, d)


def test_eigen_lower_triangle_is_ignored(trn):
    """The rotation decisions of src/eigen.rs:165 read a[i][j] with i < j: the result is a function of the upper triangle."""
    n = 9
    m = sym(n, 5)
    dirty = m.copy()
    dirty[np.tril_indices(n, -1)] = 123.0
    a = trn.SymmetricEigen.new(trn.Matrix.from_vec(n, n, m.ravel()))
    b = trn.SymmetricEigen.new(trn.Matrix.from_vec(n, n, dirty.ravel()))
    assert np.array_equal(a.eigenvalues(), b.eigenvalues())
    assert np.array_equal(a.eigenvectors().as_slice(), b.eigenvectors().as_slice())


def test_eigen_repeated_and_clustered_spectrum(trn):
    """Projector-like matrices: eigenvalues {2 (x3), -1 (x5)} from an orthogonal basis; deterministic reruns."""
    n = 8
    q, _ = np.linalg.qr(np.random.default_rng(3).standard_normal((n, n)))
    lam = np.array([2, 2, 2, -1, -1, -1, -1, -1], np.float64)
    m = (q @ np.diag(lam) @ q.T).astype(f32)
    m = ((m + m.T) / 2).astype(f32)
    a = trn.SymmetricEigen.new(trn.Matrix.from_vec(n, n, m.ravel()))
    assert np.max(np.abs(a.eigenvalues() - lam)) <= 1e-5
    b = trn.SymmetricEigen.new(trn.Matrix.from_vec(n, n, m.ravel()))
    assert np.array_equal(a.eigenvalues(), b.eigenvalues())
    assert np.array_equal(a.eigenvectors().as_slice(), b.eigenvectors().as_slice())


def test_eigen_1024_properties(trn):
    """The reference's GPU threshold (n >= 1000, src/eigen.rs:44): a 1024 x 1024 covariance-like matrix."""
    n = 1024
    rng = np.random.default_rng(7)
    x = rng.standard_normal((n, 2 * n)).astype(f32)
    m = (x @ x.T / f32(2 * n)).astype(f32)
    m = ((m + m.T) / 2).astype(f32)
    eig = trn.SymmetricEigen.new(trn.Matrix.from_vec(n, n, m.ravel()))
    vals = eig.eigenvalues().astype(np.float64)
    vecs = eig.eigenvectors().as_slice().reshape(n, n).astype(np.float64)
    frob = np.linalg.norm(m.astype(np.float64))
    assert np.all(np.diff(vals) <= 0)
    assert np.max(np.abs(vals - np.linalg.eigvalsh(m.astype(np.float64))[::-1])) <= 2e-6 * frob
    assert np.max(np.abs(vecs.T @ vecs - np.eye(n))) <= 1e-5 * n ** 0.5
    assert np.max(np.abs(m.astype(np.float64) @ vecs - vecs * vals[None, :])) <= 4e-6 * frob
    assert abs(vals.sum() - np.trace(m.astype(np.float64))) <= 1e-5 * frob * n ** 0.5
