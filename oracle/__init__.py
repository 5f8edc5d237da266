"""ctypes loader for the CPU parity oracle (oracle/trueno_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference leg.  The product package (trueno_b200/) must never import this.

The functions mirror the reference's public API semantics (paiml/trueno src/vector.rs,
src/matrix.rs) on numpy float32 arrays and raise the same error taxonomy (src/error.rs:8-41).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

SCALAR, AVX2, AVX512 = 0, 1, 2


class OracleError(Exception):
    """Carries the TruenoError variant name and message the reference would produce."""

    def __init__(self, variant: str, message: str, expected: int | None = None, actual: int | None = None):
        super().__init__(f"{variant}: {message}")
        self.variant, self.message, self.expected, self.actual = variant, message, expected, actual


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "trueno_oracle.c"))
    ):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _LIB_PATH


_f32p = C.POINTER(C.c_float)
_f64p = C.POINTER(C.c_double)
_u64p = C.POINTER(C.c_uint64)


def _p(a: np.ndarray):
    return a.ctypes.data_as(_f32p)


def _f32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32).reshape(-1)


class Oracle:
    def __init__(self):
        self.lib = C.CDLL(build())
        L = self.lib
        L.orc_last_message.restype = C.c_char_p
        for name in ("scalar_dot", "avx2_dot", "avx512_dot"):
            f = getattr(L, "orc_" + name); f.restype = C.c_float; f.argtypes = [_f32p, _f32p, C.c_size_t]
        for name in ("scalar_sum", "avx2_sum", "avx512_sum", "scalar_max", "scalar_min", "avx2_max", "avx2_min",
                     "scalar_norm_l2", "avx2_norm_l2"):
            f = getattr(L, "orc_" + name); f.restype = C.c_float; f.argtypes = [_f32p, C.c_size_t]
        for name in ("scalar_argmax", "scalar_argmin", "avx2_argmax", "avx2_argmin"):
            f = getattr(L, "orc_" + name); f.restype = C.c_uint64; f.argtypes = [_f32p, C.c_size_t]
        for name in ("scalar_sigmoid", "scalar_gelu", "avx2_sigmoid", "avx2_gelu", "avx2_exp"):
            f = getattr(L, "orc_" + name); f.restype = None; f.argtypes = [_f32p, _f32p, C.c_size_t]
        for name in ("scalar_add", "scalar_mul", "avx2_add", "avx2_mul"):
            f = getattr(L, "orc_" + name); f.restype = None; f.argtypes = [_f32p, _f32p, _f32p, C.c_size_t]
        L.orc_vector_dot.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, C.c_int, _f32p]
        for name in ("sum", "max", "min", "norm_l2"):
            getattr(L, "orc_vector_" + name).argtypes = [_f32p, C.c_size_t, C.c_int, _f32p]
        for name in ("argmax", "argmin"):
            getattr(L, "orc_vector_" + name).argtypes = [_f32p, C.c_size_t, C.c_int, _u64p]
        for name in ("add", "mul"):
            getattr(L, "orc_vector_" + name).argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, _f32p]
        for name in ("sigmoid", "gelu", "softmax", "log_softmax"):
            getattr(L, "orc_vector_" + name).argtypes = [_f32p, C.c_size_t, C.c_int, _f32p]
        L.orc_softmax_rows.argtypes = [_f32p, _f32p, C.c_size_t, C.c_size_t, C.c_int, C.c_int]
        L.orc_transpose.restype = None
        L.orc_transpose.argtypes = [_f32p, _f32p, C.c_size_t, C.c_size_t]
        L.orc_matmul_naive.restype = None
        L.orc_matmul_naive.argtypes = [_f32p, _f32p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t]
        L.orc_matmul_simd.restype = None
        L.orc_matmul_simd.argtypes = [_f32p, _f32p, _f32p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]
        L.orc_matmul.argtypes = [_f32p, C.c_size_t, C.c_size_t, _f32p, C.c_size_t, C.c_size_t, _f32p, C.c_int]
        L.orc_batched_matmul.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, _f32p] + [C.c_size_t] * 4 + [C.c_int]
        L.orc_batched_matmul_4d.argtypes = [_f32p, C.c_size_t, _f32p, C.c_size_t, _f32p] + [C.c_size_t] * 5 + [C.c_int]
        L.orc_matvec.argtypes = [_f32p, C.c_size_t, C.c_size_t, _f32p, C.c_size_t, _f32p, C.c_int]
        L.orc_map_parallel.restype = None
        L.orc_map_parallel.argtypes = [C.c_int, _f32p, _f32p, _f32p, C.c_size_t]
        L.orc_f64_sum.restype = None
        L.orc_f64_sum.argtypes = [_f32p, C.c_size_t, _f64p, _f64p]
        L.orc_f64_dot.restype = None
        L.orc_f64_dot.argtypes = [_f32p, _f32p, C.c_size_t, _f64p, _f64p]
        L.orc_f64_matmul_samples.restype = None
        L.orc_f64_matmul_samples.argtypes = [_f32p, _f32p, C.c_size_t, C.c_size_t, _u64p, _u64p, C.c_size_t, _f64p, _f64p]
        L.orc_scalar_map.restype = None
        L.orc_scalar_map.argtypes = [C.c_int, _f32p, _f32p, _f32p, C.c_float, C.c_float, _f32p, C.c_size_t]
        for name in ("scalar_sum_kahan", "scalar_norm_l1", "scalar_norm_linf"):
            f = getattr(L, "orc_" + name); f.restype = C.c_float; f.argtypes = [_f32p, C.c_size_t]
        L.orc_vector_map.restype = None
        L.orc_vector_map.argtypes = [C.c_int, _f32p, _f32p, C.c_float, C.c_float, _f32p, C.c_size_t]
        L.orc_layer_norm_simple.restype = None
        L.orc_layer_norm_simple.argtypes = [_f32p, C.c_float, C.c_float, _f32p, C.c_size_t]
        L.orc_embedding_lookup.restype = C.c_int
        L.orc_embedding_lookup.argtypes = [_f32p, C.c_size_t, C.c_size_t, _u64p, C.c_size_t, _f32p]
        L.orc_symmetric_eigen.restype = C.c_int
        L.orc_symmetric_eigen.argtypes = [_f32p, C.c_size_t, _f32p, _f32p]
        L.orc_convolve2d.restype = None
        L.orc_convolve2d.argtypes = [_f32p, C.c_size_t, C.c_size_t, _f32p, C.c_size_t, C.c_size_t, _f32p]
        L.orc_vecmat.restype = None
        L.orc_vecmat.argtypes = [_f32p, _f32p, C.c_size_t, C.c_size_t, _f32p]
        L.orc_layer_norm.restype = None
        L.orc_layer_norm.argtypes = [_f32p, _f32p, _f32p, C.c_float, _f32p, C.c_size_t]
        L.orc_last_mismatch.restype = None
        L.orc_last_mismatch.argtypes = [_u64p, _u64p]
        L.orc_set_threads.argtypes = [C.c_int]
        self.has_avx512 = bool(L.orc_has_avx512())

    # -- error mapping -------------------------------------------------------------------------
    def _check(self, status: int):
        if status == 0:
            return
        msg = self.lib.orc_last_message().decode("utf-8")
        if status == 1:
            e, a = C.c_uint64(), C.c_uint64()
            self.lib.orc_last_mismatch(C.byref(e), C.byref(a))
            raise OracleError("SizeMismatch", msg, e.value, a.value)
        raise OracleError({2: "InvalidInput", 3: "EmptyVector"}[status], msg)

    def num_threads(self) -> int:
        return int(self.lib.orc_num_threads())

    def set_threads(self, t: int):
        self.lib.orc_set_threads(int(t))

    # -- Vector API (src/vector.rs) ------------------------------------------------------------
    def dot(self, a, b, backend=AVX2) -> np.float32:
        a, b = _f32(a), _f32(b)
        out = C.c_float()
        self._check(self.lib.orc_vector_dot(_p(a), a.size, _p(b), b.size, backend, C.byref(out)))
        return np.float32(out.value)

    def _red(self, name, a, backend):
        a = _f32(a)
        out = C.c_float()
        self._check(getattr(self.lib, "orc_vector_" + name)(_p(a), a.size, backend, C.byref(out)))
        return np.float32(out.value)

    def sum(self, a, backend=AVX2): return self._red("sum", a, backend)
    def max(self, a, backend=AVX2): return self._red("max", a, backend)
    def min(self, a, backend=AVX2): return self._red("min", a, backend)
    def norm_l2(self, a, backend=AVX2): return self._red("norm_l2", a, backend)

    def _arg(self, name, a, backend):
        a = _f32(a)
        out = C.c_uint64()
        self._check(getattr(self.lib, "orc_vector_" + name)(_p(a), a.size, backend, C.byref(out)))
        return int(out.value)

    def argmax(self, a, backend=SCALAR): return self._arg("argmax", a, backend)
    def argmin(self, a, backend=SCALAR): return self._arg("argmin", a, backend)

    def _bin(self, name, a, b):
        a, b = _f32(a), _f32(b)
        out = np.empty(a.size, np.float32)
        self._check(getattr(self.lib, "orc_vector_" + name)(_p(a), a.size, _p(b), b.size, _p(out)))
        return out

    def add(self, a, b): return self._bin("add", a, b)
    def mul(self, a, b): return self._bin("mul", a, b)

    def _map(self, name, a, backend):
        a = _f32(a)
        out = np.empty(a.size, np.float32)
        self._check(getattr(self.lib, "orc_vector_" + name)(_p(a), a.size, backend, _p(out)))
        return out

    def sigmoid(self, a, backend=SCALAR): return self._map("sigmoid", a, backend)
    def gelu(self, a, backend=SCALAR): return self._map("gelu", a, backend)
    def softmax(self, a, backend=AVX2): return self._map("softmax", a, backend)
    def log_softmax(self, a, backend=AVX2): return self._map("log_softmax", a, backend)

    # ---- remaining VectorBackend surface, scalar backend (src/backends/scalar.rs) ----
    MAP_OPS = {"sub": 0, "div": 1, "scale": 2, "abs": 3, "clamp": 4, "lerp": 5, "fma": 6, "relu": 7, "exp": 8, "swish": 9,
               "tanh": 10, "sqrt": 11, "recip": 12, "ln": 13, "log2": 14, "log10": 15, "sin": 16, "cos": 17, "tan": 18,
               "floor": 19, "ceil": 20, "round": 21}

    def scalar_map(self, op: str, a, b=None, c=None, p0: float = 0.0, p1: float = 0.0) -> np.ndarray:
        a = _f32(a)
        b = _f32(b) if b is not None else a
        c = _f32(c) if c is not None else a
        out = np.empty(a.size, np.float32)
        self.lib.orc_scalar_map(self.MAP_OPS[op], _p(a), _p(b), _p(c), p0, p1, _p(out), a.size)
        return out

    # ---- the rest of Vector's element-wise / statistics API (scalar closures in src/vector.rs) ----
    VMAP_OPS = {"neg": 0, "signum": 1, "trunc": 2, "fract": 3, "sinh": 4, "cosh": 5, "asin": 6, "acos": 7, "atan": 8,
                "asinh": 9, "acosh": 10, "atanh": 11, "hardswish": 12, "mish": 13, "selu": 14, "leaky_relu": 15, "elu": 16,
                "pow": 17, "clip": 18, "minimum": 19, "maximum": 20, "copysign": 21, "affine": 22}

    def vector_map(self, op: str, a, b=None, p0: float = 0.0, p1: float = 0.0) -> np.ndarray:
        a = _f32(a)
        b = _f32(b) if b is not None else a
        out = np.empty(a.size, np.float32)
        self.lib.orc_vector_map(self.VMAP_OPS[op], _p(a), _p(b), p0, p1, _p(out), a.size)
        return out

    # statistics exactly as Vector composes them (src/vector.rs:898-1290) from its own sum / dot / min / max
    def mean(self, a, backend=AVX2):
        a = _f32(a)
        if a.size == 0:
            raise OracleError("EmptyVector", "Empty vector")
        return np.float32(np.float32(self.sum(a, backend)) / np.float32(a.size))

    def variance(self, a, backend=AVX2):   # src/vector.rs:983-1015: E[x^2] - mean^2
        a = _f32(a)
        m = self.mean(a, backend)
        ex2 = np.float32(np.float32(self.dot(a, a, backend)) / np.float32(a.size))
        return np.float32(ex2 - np.float32(m * m))

    def stddev(self, a, backend=AVX2): return np.float32(np.sqrt(self.variance(a, backend)))

    def covariance(self, a, b, backend=AVX2):   # :1063-1085
        a, b = _f32(a), _f32(b)
        if a.size == 0:
            raise OracleError("EmptyVector", "Empty vector")
        if a.size != b.size:
            raise OracleError("SizeMismatch", f"Size mismatch: expected {a.size}, got {b.size}", a.size, b.size)
        mx, my = self.mean(a, backend), self.mean(b, backend)
        mxy = np.float32(np.float32(self.dot(a, b, backend)) / np.float32(a.size))
        return np.float32(mxy - np.float32(mx * my))

    def correlation(self, a, b, backend=AVX2):   # :1119-1135
        cov = self.covariance(a, b, backend)
        sx, sy = self.stddev(a, backend), self.stddev(b, backend)
        if abs(sx) < 1e-10 or abs(sy) < 1e-10:
            raise OracleError("DivisionByZero", "Division by zero")
        return np.float32(min(max(np.float32(cov / np.float32(sx * sy)), np.float32(-1)), np.float32(1)))

    def zscore(self, a, backend=AVX2):   # :1180-1203
        a = _f32(a)
        m, sd = self.mean(a, backend), self.stddev(a, backend)
        if abs(sd) < 1e-10:
            raise OracleError("DivisionByZero", "Division by zero")
        return self.vector_map("affine", a, p0=float(m), p1=float(np.float32(1) / sd))

    def minmax_normalize(self, a):   # :1248-1272
        a = _f32(a)
        if a.size == 0:
            raise OracleError("EmptyVector", "Empty vector")
        lo, hi = np.float32(self.min(a)), np.float32(self.max(a))
        rng = np.float32(hi - lo)
        if abs(rng) < 1e-10:
            raise OracleError("DivisionByZero", "Division by zero")
        return self.vector_map("affine", a, p0=float(lo), p1=float(np.float32(1) / rng))

    def layer_norm_simple(self, a, eps, backend=AVX2):   # :1386-1412
        a = _f32(a)
        if a.size == 0:
            raise OracleError("EmptyVector", "Empty vector")
        out = np.empty(a.size, np.float32)
        self.lib.orc_layer_norm_simple(_p(a), float(self.sum(a, backend)), eps, _p(out), a.size)
        return out

    def sum_kahan(self, a): a = _f32(a); return np.float32(self.lib.orc_scalar_sum_kahan(_p(a), a.size))
    def norm_l1(self, a): a = _f32(a); return np.float32(self.lib.orc_scalar_norm_l1(_p(a), a.size))
    def norm_linf(self, a): a = _f32(a); return np.float32(self.lib.orc_scalar_norm_linf(_p(a), a.size))

    def convolve2d(self, A, rows, cols, K, kr, kc):
        A, K = _f32(A), _f32(K)
        out = np.empty((rows - kr + 1) * (cols - kc + 1), np.float32)
        self.lib.orc_convolve2d(_p(A), rows, cols, _p(K), kr, kc, _p(out))
        return out.reshape(rows - kr + 1, cols - kc + 1)

    def symmetric_eigen(self, A, rows, cols):
        """SymmetricEigen::new (src/eigen.rs:108-141) on the CPU Jacobi path (src/eigen.rs:143-221): returns
        (eigenvalues descending, eigenvectors as the COLUMNS of a rows x rows matrix)."""
        if rows != cols:
            raise OracleError("InvalidInput", f"Matrix must be square for eigendecomposition, got {rows}x{cols}")
        if rows == 0:
            raise OracleError("InvalidInput", "Cannot compute eigendecomposition of empty matrix")
        A = _f32(A)
        vals = np.empty(rows, np.float32)
        vecs = np.empty(rows * rows, np.float32)
        if self.lib.orc_symmetric_eigen(_p(A), rows, _p(vals), _p(vecs)) != 0:
            raise OracleError("InvalidInput", "Jacobi algorithm failed to converge after 50 sweeps")
        return vals, vecs.reshape(rows, rows)

    def embedding_lookup(self, table, rows, cols, indices):
        """Matrix::embedding_lookup (src/matrix.rs:2008-2041)"""
        table = _f32(table)
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64).reshape(-1))
        out = np.empty(idx.size * cols, np.float32)
        self._check(self.lib.orc_embedding_lookup(_p(table), rows, cols, idx.ctypes.data_as(_u64p), idx.size, _p(out)))
        return out.reshape(idx.size, cols)

    def vecmat(self, v, A, rows, cols):
        v, A = _f32(v), _f32(A)
        out = np.empty(cols, np.float32)
        self.lib.orc_vecmat(_p(v), _p(A), rows, cols, _p(out))
        return out

    def layer_norm(self, x, gamma, beta, eps):
        x, gamma, beta = _f32(x), _f32(gamma), _f32(beta)
        out = np.empty(x.size, np.float32)
        self.lib.orc_layer_norm(_p(x), _p(gamma), _p(beta), eps, _p(out), x.size)
        return out

    def exp_avx2(self, a):
        a = _f32(a)
        out = np.empty(a.size, np.float32)
        self.lib.orc_avx2_exp(_p(a), _p(out), a.size)
        return out

    def softmax_rows(self, a, rows, cols, log=False, backend=AVX2):
        a = _f32(a)
        out = np.empty(a.size, np.float32)
        self._check(self.lib.orc_softmax_rows(_p(a), _p(out), rows, cols, backend, int(log)))
        return out.reshape(rows, cols)

    def map_parallel(self, op: str, a, b=None):
        a = _f32(a)
        b = a if b is None else _f32(b)
        out = np.empty(a.size, np.float32)
        self.lib.orc_map_parallel({"add": 0, "mul": 1, "sigmoid": 2, "gelu": 3}[op], _p(a), _p(b), _p(out), a.size)
        return out

    # -- Matrix API (src/matrix.rs) ------------------------------------------------------------
    def matmul(self, A, a_shape, B, b_shape, parallel=False):
        A, B = _f32(A), _f32(B)
        out = np.empty(a_shape[0] * b_shape[1], np.float32)
        self._check(self.lib.orc_matmul(_p(A), a_shape[0], a_shape[1], _p(B), b_shape[0], b_shape[1], _p(out), int(parallel)))
        return out.reshape(a_shape[0], b_shape[1])

    def matmul_naive(self, A, B, m, k, n):
        A, B = _f32(A), _f32(B)
        out = np.empty(m * n, np.float32)
        self.lib.orc_matmul_naive(_p(A), _p(B), _p(out), m, k, n)
        return out.reshape(m, n)

    def matmul_simd(self, A, B, m, k, n, parallel=False):
        A, B = _f32(A), _f32(B)
        out = np.empty(m * n, np.float32)
        self.lib.orc_matmul_simd(_p(A), _p(B), _p(out), m, k, n, int(parallel))
        return out.reshape(m, n)

    def transpose(self, A, rows, cols):
        A = _f32(A)
        out = np.empty(A.size, np.float32)
        self.lib.orc_transpose(_p(A), _p(out), rows, cols)
        return out.reshape(cols, rows)

    def batched_matmul(self, A, B, batch, m, k, n, parallel=False):
        A, B = _f32(A), _f32(B)
        out = np.empty(batch * m * n, np.float32)
        self._check(self.lib.orc_batched_matmul(_p(A), A.size, _p(B), B.size, _p(out), batch, m, k, n, int(parallel)))
        return out

    def batched_matmul_4d(self, A, B, batch, heads, m, k, n, parallel=False):
        A, B = _f32(A), _f32(B)
        out = np.empty(batch * heads * m * n, np.float32)
        self._check(self.lib.orc_batched_matmul_4d(_p(A), A.size, _p(B), B.size, _p(out), batch, heads, m, k, n, int(parallel)))
        return out

    def attention(self, q, k, v, heads, seq, d, scale=None, causal=False):
        """TEST ORACLE for the fused attention: the composition the reference spells with its own CPU operators —
        Matrix::transpose (src/matrix.rs:1590) + Matrix::matmul (src/matrix.rs:285) for Q K^T (the "attention
        pattern" of src/matrix.rs:3985), ScalarBackend::scale, Vector::softmax per row (src/vector.rs:1516), matmul
        with V — with trueno-gpu's AttentionKernel conventions (trueno-gpu/src/kernels/attention.rs:46-112):
        layout [heads][seq][d], scale = 1/sqrt(d) by default, causal = keys after the query are masked."""
        q = _f32(q).reshape(heads, seq, d); k = _f32(k).reshape(heads, seq, d); v = _f32(v).reshape(heads, seq, d)
        if scale is None:
            scale = np.float32(1.0) / np.sqrt(np.float32(d))
        scale = np.float32(scale)
        out = np.empty((heads, seq, d), np.float32)
        for h in range(heads):
            kt = self.transpose(k[h].ravel(), seq, d)
            s = self.matmul(q[h].ravel(), (seq, d), kt, (d, seq)).reshape(seq, seq)
            s = self.scalar_map("scale", s.ravel(), p0=float(scale)).reshape(seq, seq)
            if causal:
                s = np.where(np.arange(seq)[None, :] > np.arange(seq)[:, None], np.float32(-np.inf), s).astype(np.float32)
            p = self.softmax_rows(s.ravel(), seq, seq)
            out[h] = self.matmul(p, (seq, seq), v[h].ravel(), (seq, d)).reshape(seq, d)
        return out.ravel()

    def matvec(self, A, rows, cols, v, parallel=False):
        A, v = _f32(A), _f32(v)
        out = np.empty(rows, np.float32)
        self._check(self.lib.orc_matvec(_p(A), rows, cols, _p(v), v.size, _p(out), int(parallel)))
        return out

    # -- f64 yardsticks ------------------------------------------------------------------------
    def f64_sum(self, a):
        a = _f32(a)
        s, t = C.c_double(), C.c_double()
        self.lib.orc_f64_sum(_p(a), a.size, C.byref(s), C.byref(t))
        return s.value, t.value

    def f64_dot(self, a, b):
        a, b = _f32(a), _f32(b)
        s, t = C.c_double(), C.c_double()
        self.lib.orc_f64_dot(_p(a), _p(b), a.size, C.byref(s), C.byref(t))
        return s.value, t.value

    def f64_matmul_samples(self, A, B, k, n, rows, cols):
        A, B = _f32(A), _f32(B)
        rows = np.ascontiguousarray(rows, np.uint64); cols = np.ascontiguousarray(cols, np.uint64)
        out = np.empty(rows.size, np.float64); oabs = np.empty(rows.size, np.float64)
        self.lib.orc_f64_matmul_samples(_p(A), _p(B), k, n, rows.ctypes.data_as(_u64p), cols.ctypes.data_as(_u64p),
                                        rows.size, out.ctypes.data_as(_f64p), oabs.ctypes.data_as(_f64p))
        return out, oabs


_singleton: Oracle | None = None


def get() -> Oracle:
    global _singleton
    if _singleton is None:
        _singleton = Oracle()
    return _singleton
