"""The oracle's restatement of the rest of Vector's element-wise / statistics API (oracle.vector_map and the composed
statistics) against the reference's own known-answer tests (tests/vector_api_kats.py)."""
import numpy as np
import pytest

import vector_api_kats

f32 = np.float32


class _Err(Exception):
    def __init__(self, variant, msg=""):
        super().__init__(msg)
        self.variant = variant


def _oracle_call(o):
    import oracle as O

    def call(op, *vecs, p=()):
        a = np.asarray(vecs[0], f32)
        try:
            if op in ("leaky_relu", "elu"):   # validation as in src/vector.rs:1981-1991, :2086-2096
                if a.size == 0:
                    raise _Err("EmptyVector")
                if op == "leaky_relu" and not (0.0 <= p[0] < 1.0):
                    raise _Err("InvalidInput", f"negative_slope must be in [0.0, 1.0), got {p[0]}")
                if op == "elu" and p[0] <= 0:
                    raise _Err("InvalidInput", f"alpha must be > 0, got {p[0]}")
                return o.vector_map(op, a, p0=p[0])
            if op in ("hardswish", "mish", "selu"):
                if a.size == 0:
                    raise _Err("EmptyVector")
                return o.vector_map(op, a)
            if op == "clip":
                if p[0] > p[1]:
                    raise _Err("InvalidInput", f"min_val ({p[0]:g}) must be <= max_val ({p[1]:g})")
                return o.vector_map(op, a, p0=p[0], p1=p[1])
            if op == "pow":
                return o.vector_map(op, a, p0=p[0])
            if op in ("minimum", "maximum", "copysign"):
                b = np.asarray(vecs[1], f32)
                if a.size != b.size:
                    raise _Err("SizeMismatch")
                return o.vector_map(op, a, b)
            if op == "sum_of_squares":
                return f32(0) if a.size == 0 else o.dot(a, a)
            if op in ("covariance", "correlation"):
                return getattr(o, op)(a, np.asarray(vecs[1], f32))
            if op in ("zscore", "minmax_normalize"):
                if a.size == 0:
                    raise _Err("EmptyVector")
                return getattr(o, op)(a)
            if op == "layer_norm_simple":
                return o.layer_norm_simple(a, p[0])
            return o.vector_map(op, a)
        except O.OracleError as e:
            raise _Err(e.variant, str(e)) from None
    return call


def test_oracle_vector_api_kats(oracle):
    vector_api_kats.run(_oracle_call(oracle))


def test_oracle_vector_map_against_numpy(oracle):
    """the oracle's C closures against numpy's f32 functions on a seeded sample (sanity of the op table)"""
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(4097) * 3).astype(f32)
    u = rng.uniform(-0.99, 0.99, 4097).astype(f32)
    ulp = lambda v: np.spacing(np.abs(v).astype(f32)).astype(np.float64)
    for op, arg, fn in (("sinh", x, np.sinh), ("cosh", x, np.cosh), ("asin", u, np.arcsin), ("acos", u, np.arccos),
                        ("atan", x, np.arctan), ("asinh", x, np.arcsinh), ("acosh", np.abs(x) + 1, np.arccosh),
                        ("atanh", u, np.arctanh), ("trunc", x, np.trunc), ("neg", x, np.negative)):
        want = fn(arg.astype(np.float64))
        got = oracle.vector_map(op, arg).astype(np.float64)
        assert np.all(np.abs(got - want) <= 2 * ulp(want) + 1e-45), op
    assert np.array_equal(oracle.vector_map("fract", x), x - np.trunc(x))
    assert np.array_equal(oracle.vector_map("copysign", x, u), np.copysign(x, u))
    assert np.array_equal(oracle.vector_map("minimum", x, u), np.fmin(x, u))
    assert np.array_equal(oracle.vector_map("maximum", x, u), np.fmax(x, u))
    hs = np.where(x <= -3, f32(0), np.where(x >= 3, x, (x * (x + f32(3))) / f32(6))).astype(f32)
    assert np.array_equal(oracle.vector_map("hardswish", x), hs)
    assert np.array_equal(oracle.vector_map("leaky_relu", x, p0=0.01), np.where(x > 0, x, f32(0.01) * x).astype(f32))


def test_oracle_embedding_lookup_kats(oracle):
    """src/matrix.rs:3727-3848"""
    import oracle as O
    t = np.arange(1, 13, dtype=f32)
    r = oracle.embedding_lookup(t, 4, 3, [1, 3, 0])
    assert np.array_equal(r, np.array([[4, 5, 6], [10, 11, 12], [1, 2, 3]], f32))
    assert np.array_equal(oracle.embedding_lookup([1, 2, 3, 4, 5, 6], 3, 2, [1]), np.array([[3, 4]], f32))
    r = oracle.embedding_lookup([1, 2, 3, 4, 5, 6], 2, 3, [0, 0, 1, 0])
    assert r.shape == (4, 3) and np.array_equal(r[0], r[1]) and np.array_equal(r[0], r[3])
    assert oracle.embedding_lookup([1, 2, 3, 4, 5, 6], 3, 2, []).shape == (0, 2)
    with pytest.raises(O.OracleError) as e:
        oracle.embedding_lookup([1, 2, 3, 4, 5, 6], 3, 2, [0, 5, 1])
    assert e.value.variant == "InvalidInput" and "Index 5 at position 1 is out of bounds for embedding table with 3 rows" in str(e.value)
    big = np.arange(1000 * 256, dtype=f32)
    r = oracle.embedding_lookup(big, 1000, 256, [0, 500, 999, 42, 100])
    assert r.shape == (5, 256) and r[0, 0] == 0 and r[1, 0] == 500 * 256 and r[2, 0] == 999 * 256
