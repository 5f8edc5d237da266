"""Oracle pin for the remaining VectorBackend surface: the scalar-backend restatement in
oracle/trueno_oracle.c against the exact-value KATs of the reference's own tests (tests/kats.py)."""
import numpy as np
import pytest

import kats

f32 = np.float32


@pytest.mark.parametrize("kat", kats.MAP_EXT_KATS, ids=[f"{k[0]}-{i}" for i, k in enumerate(kats.MAP_EXT_KATS)])
def test_oracle_map_kats(oracle, kat):
    op, inputs, params, expected, tol, _ = kat
    args = [np.asarray(x, f32) for x in inputs]
    p0 = params[0] if len(params) > 0 else 0.0
    p1 = params[1] if len(params) > 1 else 0.0
    got = oracle.scalar_map(op, args[0], args[1] if len(args) > 1 else None, args[2] if len(args) > 2 else None, p0, p1)
    want = np.asarray(expected, f32)
    if tol == 0:
        assert np.array_equal(got, want), (op, got, want)
    else:
        assert np.max(np.abs(got - want)) <= tol


@pytest.mark.parametrize("kat", [k for k in kats.REDUCE_EXT_KATS if k[0] in ("sum_kahan", "norm_l1", "norm_linf")],
                         ids=lambda k: k[0])
def test_oracle_reduce_kats(oracle, kat):
    op, v, expected, tol, _ = kat
    got = float(getattr(oracle, op)(np.asarray(v, f32)))
    assert abs(got - expected) <= tol


def test_oracle_semantics_corner_cases(oracle):
    nan = np.nan
    # relu: NaN and -0.0 -> +0.0 (strict `val > 0.0`, src/backends/scalar.rs:293)
    r = oracle.scalar_map("relu", [nan, -0.0, 3.0])
    assert r[0] == 0 and not np.signbit(r[1]) and r[2] == 3
    # clamp: f32::max / min ignore a NaN operand -> NaN clamps to min
    assert oracle.scalar_map("clamp", [nan], p0=1.0, p1=2.0)[0] == 1.0
    # norm_linf: NaN never wins
    assert float(oracle.norm_linf(np.array([1, nan, -3], f32))) == 3.0
    # fma is NOT fused in the scalar backend: a*b rounds before the add
    a, b, c = f32(1 + 2 ** -12), f32(1 + 2 ** -12), f32(-1)
    assert oracle.scalar_map("fma", [a], [b], [c])[0] == f32(f32(a * b) + c)
    # Kahan keeps what the plain left-to-right sum loses
    v = np.array([1e8] + [1.0] * 1000, f32)
    assert float(oracle.sum_kahan(v)) == 1e8 + 1000


def test_oracle_vecmat_and_layer_norm_kats(oracle):
    # src/matrix.rs:3543-3555: [1,2] x [[1,2,3],[4,5,6]] = [9, 12, 15]
    assert oracle.vecmat([1, 2], [1, 2, 3, 4, 5, 6], 2, 3).tolist() == [9.0, 12.0, 15.0]
    # src/vector.rs:7656-7703: normalised output has mean ~0 / variance ~1, then scale 2 / shift 1
    y = oracle.layer_norm([1, 2, 3, 4], [1] * 4, [0] * 4, 1e-5)
    assert abs(y.mean()) < 1e-5 and abs(y.var() - 1) < 1e-3
    y = oracle.layer_norm([1, 2, 3, 4], [2] * 4, [1] * 4, 1e-5)
    assert abs(y.mean() - 1) < 1e-3 and abs(y.std() - 2) < 1e-3
    assert np.all(np.abs(oracle.layer_norm([5] * 4, [1] * 4, [0] * 4, 1e-5)) < 1e-3)   # :7745
    assert abs(oracle.layer_norm([42], [1], [0], 1e-5)[0]) < 1e-3                       # :7760


def test_oracle_convolve2d_kats(oracle):
    # src/matrix.rs:3624-3638: 1x1 identity kernel preserves the input
    inp = np.arange(1, 10, dtype=f32)
    assert np.array_equal(oracle.convolve2d(inp, 3, 3, [1.0], 1, 1).reshape(-1), inp)
    # src/matrix.rs:3676-3712: 3x3 averaging of a centred 9 -> centre 1.0
    img = np.zeros(25, f32); img[12] = 9
    out = oracle.convolve2d(img, 5, 5, np.full(9, f32(1.0) / f32(9.0), f32), 3, 3)
    assert out.shape == (3, 3) and abs(out[1, 1] - 1.0) < 1e-5
    # src/matrix.rs:3641-3673: horizontal edge kernel on the 4x4 plateau -> 2x2
    edge = oracle.convolve2d([1, 1, 1, 1, 1, 2, 2, 1, 1, 2, 2, 1, 1, 1, 1, 1], 4, 4, [-1, -1, -1, 0, 0, 0, 1, 1, 1], 3, 3)
    assert edge.shape == (2, 2) and edge.tolist() == [[2.0, 2.0], [-2.0, -2.0]]   # shape pinned by the reference; values by hand: -(1+1+1) + (1+2+2)
