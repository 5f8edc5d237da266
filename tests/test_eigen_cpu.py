"""The SymmetricEigen oracle (oracle/trueno_oracle.c::orc_symmetric_eigen, restating src/eigen.rs:143-306) pinned to
the reference's own tests (src/eigen.rs:526-870) — CPU only."""
import numpy as np
import pytest

from eigen_kats import COV_MATRIX, EIGEN_KATS, ORTHO_MATRIX

f32 = np.float32


@pytest.mark.parametrize("kat", EIGEN_KATS, ids=[k[3] for k in EIGEN_KATS])
def test_oracle_eigen_kats(oracle, kat):
    rows, want, tol, _ = kat
    n = len(rows)
    vals, vecs = oracle.symmetric_eigen(np.asarray(rows, f32).ravel(), n, n)
    assert np.all(np.diff(vals) <= 0)                                # descending
    assert np.max(np.abs(vals - np.asarray(want, f32))) < tol
    a = np.asarray(rows, np.float64)
    for i in range(n):                                               # A v = lambda v  (src/eigen.rs:644-668, 1e-4)
        assert np.max(np.abs(a @ vecs[:, i] - vals[i] * vecs[:, i])) < 1e-4


def test_oracle_eigen_orthogonal_and_covariance(oracle):
    _, v = oracle.symmetric_eigen(np.asarray(ORTHO_MATRIX, f32).ravel(), 3, 3)
    assert np.max(np.abs(v.T @ v - np.eye(3))) < 1e-4                # src/eigen.rs:591-616
    vals, _ = oracle.symmetric_eigen(np.asarray(COV_MATRIX, f32).ravel(), 2, 2)
    assert vals[0] > 5.0 and abs(vals[1]) < 0.1                      # src/eigen.rs:739-750


def test_oracle_eigen_errors(oracle):
    import oracle as orc
    with pytest.raises(orc.OracleError) as e:                        # src/eigen.rs:672-682
        oracle.symmetric_eigen(np.arange(6, dtype=f32), 2, 3)
    assert "Matrix must be square for eigendecomposition, got 2x3" in str(e.value)
    with pytest.raises(orc.OracleError) as e:                        # src/eigen.rs:686-690
        oracle.symmetric_eigen(np.zeros(0, f32), 0, 0)
    assert "Cannot compute eigendecomposition of empty matrix" in str(e.value)


@pytest.mark.parametrize("n", [2, 3, 5, 8, 33, 100])
def test_oracle_eigen_vs_lapack(oracle, n):
    """src/eigen.rs:805-870 (proptests: descending, count, reconstruction) + the f64 LAPACK spectrum."""
    rng = np.random.default_rng(n)
    m = rng.standard_normal((n, n)).astype(f32)
    m = ((m + m.T) / 2).astype(f32)
    vals, vecs = oracle.symmetric_eigen(m.ravel(), n, n)
    frob = np.linalg.norm(m.astype(np.float64))
    assert vals.size == n and np.all(np.diff(vals) <= 0)
    assert np.max(np.abs(vals - np.linalg.eigvalsh(m.astype(np.float64))[::-1])) <= 2e-6 * frob
    assert np.max(np.abs(vecs.T @ vecs - np.eye(n))) <= 1e-5 * max(1, n ** 0.5)
    assert np.max(np.abs(vecs @ np.diag(vals) @ vecs.T - m)) <= 4e-6 * frob
