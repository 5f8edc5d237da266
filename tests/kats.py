"""Known-answer tests and seeded fixtures lifted from the reference's own test-suite for the
hot path.  Each entry cites the reference file:line that pins it.  The same
tables drive tests/test_oracle_kat.py (CPU oracle) and tests/test_parity_gpu.py (CUDA path)."""
import numpy as np

f32 = np.float32


def arr(*xs):
    return np.array(xs, dtype=np.float32)


# (name, op, args, expected, abs_tol, citation)
REDUCTION_KATS = [
    ("dot_basic", "dot", (arr(1, 2, 3), arr(4, 5, 6)), 32.0, 0.0, "src/vector.rs:4689-4694"),
    ("dot_9", "dot", (np.arange(1, 10, dtype=f32), np.arange(1, 10, dtype=f32)[::-1].copy()), 165.0, 0.0,
     "src/backends/avx2.rs:1654-1665"),
    ("sum_basic", "sum", (arr(1, 2, 3, 4),), 10.0, 0.0, "src/vector.rs:4714"),
    ("sum_empty", "sum", (arr(),), 0.0, 0.0, "src/vector.rs:4720"),
    ("sum_single", "sum", (arr(42),), 42.0, 0.0, "src/vector.rs:4727"),
    ("sum_9", "sum", (np.arange(1, 10, dtype=f32),), 45.0, 0.0, "src/backends/avx2.rs:1670-1680"),
    ("max_9", "max", (arr(3, 1, 4, 1, 5, 9, 2, 6, 5),), 9.0, 0.0, "src/backends/avx2.rs:1685-1700"),
    ("min_9", "min", (arr(3, 1, 4, 1, 5, 9, 2, 6, 5),), 1.0, 0.0, "src/backends/avx2.rs:1705-1717"),
    ("norm_l2_345", "norm_l2", (arr(3, 4),), 5.0, 0.0, "src/vector.rs:2586-2588"),
    ("norm_l2_empty", "norm_l2", (arr(),), 0.0, 0.0, "src/vector.rs:2596-2598"),
]

ARG_KATS = [
    ("argmax_basic", "argmax", arr(1, 5, 3, 2), 1, "src/vector.rs:4837-4842"),
    ("argmax_negative", "argmax", arr(-5, -1, -3, -2), 1, "src/vector.rs:4844-4847"),
    ("argmax_ties_first", "argmax", arr(1, 5, 3, 5, 2), 1, "src/vector.rs:4866-4870"),
    ("argmin_basic", "argmin", arr(5, 1, 3, 2), 1, "src/vector.rs:4873-4878"),
    ("argmin_ties_first", "argmin", arr(5, 1, 3, 1, 2), 1, "src/vector.rs:4902-4906"),
]


def planted32(kind):
    """32-element vector with a planted extremum (src/vector.rs:9010-9027)."""
    v = np.arange(32, dtype=f32)
    if kind == "max":
        v[25] = 1000.0
        return v, 25
    v[18] = -500.0
    return v, 18


MATMUL_KATS = [
    # name, A(m,k), B(k,n), expected, tol, cite
    ("2x2", arr(1, 2, 3, 4).reshape(2, 2), arr(5, 6, 7, 8).reshape(2, 2), arr(19, 22, 43, 50).reshape(2, 2), 0.0,
     "src/matrix.rs:2166-2179; tests/wasm_optimization_tests.rs:363-379"),
    ("2x3_3x2", arr(1, 2, 3, 4, 5, 6).reshape(2, 3), arr(7, 8, 9, 10, 11, 12).reshape(3, 2),
     arr(58, 64, 139, 154).reshape(2, 2), 0.0, "src/matrix.rs:2182-2196"),
    ("1x1", arr(3).reshape(1, 1), arr(4).reshape(1, 1), arr(12).reshape(1, 1), 0.0, "src/matrix.rs:2215-2222"),
    ("7x8_8x5_ones", np.ones((7, 8), f32), np.ones((8, 5), f32), np.full((7, 5), 8.0, f32), 0.0,
     "src/matrix.rs:2268-2279"),
]

BATCHED_KAT = dict(
    a=np.arange(1, 13, dtype=f32), b=np.arange(1, 13, dtype=f32), batch=2, m=2, k=3, n=2,
    expected=arr(22, 28, 49, 64, 220, 244, 301, 334), cite="src/matrix.rs:3853-3889")

BATCHED4D_KAT = dict(
    a=np.arange(1, 9, dtype=f32), b=arr(1, 0, 0, 1, 1, 0, 0, 1), batch=1, heads=2, m=2, k=2, n=2,
    expected=np.arange(1, 9, dtype=f32), cite="src/matrix.rs:3948-3982")

MATVEC_KAT = dict(a=arr(1, 2, 3, 4, 5, 6), rows=2, cols=3, v=arr(1, 2, 3), expected=arr(14, 32),
                  cite="src/matrix.rs:1648-1655")


def fixture_mod(size_m, size_k, size_n, am, ad, bmul, bm, bd):
    """A[i] = (i % am)/ad ; B[i] = ((i*bmul) % bm)/bd — the family used at src/matrix.rs:2392-2714."""
    ia = np.arange(size_m * size_k, dtype=np.int64)
    ib = np.arange(size_k * size_n, dtype=np.int64)
    A = ((ia % am).astype(f32) / f32(ad)).reshape(size_m, size_k)
    B = (((ib * bmul) % bm).astype(f32) / f32(bd)).reshape(size_k, size_n)
    return A, B


class SimpleRng:
    """xorshift64 from tests/pixel_fkr.rs:132-153: state ^= state<<13; ^= >>7; ^= <<17;
    f = (state as f32 / u64::MAX as f32) * 2 - 1."""

    def __init__(self, seed):
        self.state = np.uint64(seed)

    def gen_vec(self, n):
        out = np.empty(n, dtype=f32)
        s = int(self.state)
        M = (1 << 64) - 1
        umax = f32(np.float32(M))  # u64::MAX as f32 == 2^64
        for i in range(n):
            s ^= (s << 13) & M
            s ^= s >> 7
            s ^= (s << 17) & M
            out[i] = f32(f32(np.float32(s)) / umax) * f32(2.0) - f32(1.0)
        self.state = np.uint64(s)
        return out


def splitmix_u01(seed, start, count):
    """Counter-based generator of SURVEY.md §8d: x = u01(splitmix64(seed ^ idx)), 24-bit mantissa."""
    idx = np.arange(start, start + count, dtype=np.uint64)
    z = (idx ^ np.uint64(seed)) + np.uint64(0x9E3779B97F4A7C15)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    z = z ^ (z >> np.uint64(31))
    return ((z >> np.uint64(40)).astype(np.float32) * f32(1.0 / (1 << 24))).astype(f32)


# ---- remaining VectorBackend surface: exact-value KATs of the reference's own tests -----
# (op, inputs, params, expected, tolerance (0 = exact), reference test location)
MAP_EXT_KATS = [
    ("sub", [[5, 7, 9], [1, 2, 3]], (), [4, 5, 6], 0, "src/vector.rs:4567"),
    ("div", [[10, 20, 30], [2, 4, 5]], (), [5, 5, 6], 0, "src/vector.rs:4640"),
    ("abs", [[3, -4, 5, -2]], (), [3, 4, 5, 2], 0, "src/vector.rs:5072"),
    ("scale", [[1, 2, 3, 4]], (2.0,), [2, 4, 6, 8], 0, "src/vector.rs:5108"),
    ("scale", [[1, 2, 3]], (0.0,), [0, 0, 0], 0, "src/vector.rs:5116"),
    ("clamp", [[-5, 0, 5, 10, 15]], (0.0, 10.0), [0, 0, 5, 10, 10], 0, "src/vector.rs:5151"),
    ("lerp", [[0, 10, 20], [100, 110, 120]], (0.5,), [50, 60, 70], 0, "src/vector.rs:5208"),
    ("lerp", [[1, 2, 3], [4, 5, 6]], (0.0,), [1, 2, 3], 0, "src/vector.rs:5216"),
    ("fma", [[2, 3, 4], [5, 6, 7], [1, 2, 3]], (), [11, 20, 31], 0, "src/vector.rs:5265"),
    ("sqrt", [[4, 9, 16, 25]], (), [2, 3, 4, 5], 0, "src/vector.rs:5333"),
    ("recip", [[2, 4, 5, 10]], (), [0.5, 0.25, 0.2, 0.1], 0, "src/vector.rs:5379"),
    ("floor", [[3.7, -2.3, 5.0]], (), [3, -3, 5], 0, "src/vector.rs:6689"),
    ("ceil", [[3.2, -2.7, 5.0]], (), [4, -2, 5], 0, "src/vector.rs:6731"),
    ("round", [[3.2, 3.7, -2.3, -2.8, 5.0]], (), [3, 4, -2, -3, 5], 0, "src/vector.rs:6773"),
    ("round", [[1.4, 1.5, 1.6, 2.5]], (), [1, 2, 2, 3], 0, "src/vector.rs:6780"),
    ("round", [[-1.4, -1.5, -1.6, -2.5]], (), [-1, -2, -2, -3], 0, "src/vector.rs:6787"),
    ("round", [[0.5, 1.5, 2.5, 3.5, 4.5]], (), [1, 2, 3, 4, 5], 0, "src/vector.rs:6795"),
    ("relu", [[-2, -1, 0, 1, 2]], (), [0, 0, 0, 1, 2], 0, "src/vector.rs:8018"),
    ("swish", [[-2, -1, 0, 1, 2]], (), [-0.238, -0.269, 0.0, 0.731, 1.762], 0.01, "src/vector.rs:8424"),
    ("exp", [[0, 1, 2]], (), [1.0, 2.718281828, 7.389056099], 1e-5, "src/vector.rs:5475"),
    ("tanh", [[0.0]], (), [0.0], 0, "src/vector.rs:6472"),
    ("sin", [[0.0]], (), [0.0], 0, "src/vector.rs:5813"),
    ("cos", [[0.0]], (), [1.0], 0, "src/vector.rs:5905"),
]
REDUCE_EXT_KATS = [
    ("sum_kahan", [1, 2, 3, 4], 10.0, 0, "src/vector.rs:4733"),
    ("sum_kahan", [], 0.0, 0, "src/vector.rs:4739"),
    ("sum_kahan", [42], 42.0, 0, "src/vector.rs:4744"),
    ("norm_l1", [3, -4, 5], 12.0, 1e-5, "src/vector.rs:4987"),
    ("norm_l1", [1, 2, 3, 4], 10.0, 1e-5, "src/vector.rs:4994"),
    ("norm_linf", [3, -7, 5, -2], 7.0, 1e-5, "src/vector.rs:5026"),
    ("norm_linf", [1, 2, 5, 3], 5.0, 1e-5, "src/vector.rs:5033"),
    ("mean", [1, 2, 3, 4], 2.5, 1e-5, "src/vector.rs:7238"),
    ("mean", [-2, -4, -6], -4.0, 1e-5, "src/vector.rs:7245"),
    ("variance", [1, 2, 3, 4, 5], 2.0, 1e-5, "src/vector.rs:7284"),
    ("variance", [7, 7, 7, 7], 0.0, 1e-5, "src/vector.rs:7291"),
    ("stddev", [1, 2, 3, 4, 5], 2.0 ** 0.5, 1e-5, "src/vector.rs:7335"),
]
