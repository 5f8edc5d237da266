"""Known answers of the reference's own SymmetricEigen tests (src/eigen.rs:526-795) — shared by the oracle (CPU) and
the CUDA path (GPU) tests.  (matrix rows, expected eigenvalues descending, tolerance, citation)"""
EIGEN_KATS = [
    ([[2, 1], [1, 2]], [3, 1], 1e-5, "src/eigen.rs:527-551"),
    ([[1, 0, 0], [0, 1, 0], [0, 0, 1]], [1, 1, 1], 1e-5, "src/eigen.rs:555-571"),
    ([[5, 0, 0], [0, 3, 0], [0, 0, 1]], [5, 3, 1], 1e-5, "src/eigen.rs:575-587"),
    ([[3, 1], [1, 3]], [4, 2], 1e-5, "src/eigen.rs:373-378 (doc test), :644-668"),
    ([[7]], [7], 1e-6, "src/eigen.rs:694-700"),
    ([[2, 0], [0, 1]], [2, 1], 1e-6, "src/eigen.rs:704-713"),
    ([[0, 1], [1, 0]], [1, -1], 1e-5, "src/eigen.rs:754-769"),
    ([[4, 2], [2, 4]], [6, 2], 1e-5, "src/eigen.rs:620-640"),
]
# src/eigen.rs:591-616 (orthogonality, 1e-4), :739-750 (covariance: > 5.0 and |.| < 0.1)
ORTHO_MATRIX = [[4, 2, 0], [2, 5, 3], [0, 3, 6]]
COV_MATRIX = [[2.67, 2.67], [2.67, 2.67]]
