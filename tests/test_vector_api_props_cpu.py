"""The reference's proptests for the plain element-wise maps (src/vector.rs, `test_*_correctness`, `*_inverse`,
`*_identity`, `*_odd_function` ...: 100 cases each, same generators and tolerances) restated with hypothesis against the
ORACLE's restatement (oracle.vector_map) — they pin the oracle on CPU; the CUDA path is held to the same properties in
tests/test_properties_gpu.py and to the oracle in tests/test_vector_api_gpu.py."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st
from hypothesis.extra import numpy as hnp

f32 = np.float32
CASES = settings(max_examples=100, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.function_scoped_fixture])


def vec(lo, hi, max_len=100):
    return hnp.arrays(f32, st.integers(1, max_len), elements=st.floats(lo, hi, width=32, allow_nan=False))


@CASES
@given(a=vec(-5.0, 5.0))
def test_hyperbolic_identities_and_inverses(oracle, a):
    s, c = oracle.vector_map("sinh", a), oracle.vector_map("cosh", a)
    ident = c * c - s * s                                                      # test_cosh_sinh_identity
    m2 = np.maximum(np.abs(c), np.abs(s)) ** 2
    assert np.all(np.abs(ident - 1) < np.where(m2 > 1, m2 * 1e-4, 1e-5))
    assert np.array_equal(oracle.vector_map("sinh", -a), -s)                   # odd / even functions
    assert np.array_equal(oracle.vector_map("cosh", -a), c)
    assert np.all(np.abs(oracle.vector_map("asinh", s) - a) < 1e-5 * np.maximum(1, np.abs(a)) + 2e-5)   # test_asinh_sinh_inverse
    pos = np.abs(a) + f32(0.1)
    assert np.all(np.abs(oracle.vector_map("acosh", oracle.vector_map("cosh", pos)) - pos) < 2e-5 * np.maximum(1, pos) + 1e-4)   # test_acosh_cosh_inverse
    assert np.all(c >= 1) and np.all(oracle.vector_map("acosh", c) >= 0)       # test_acosh_range


@CASES
@given(a=vec(-3.5, 3.5))
def test_atanh_tanh_inverse_and_odd(oracle, a):
    t = oracle.scalar_map("tanh", a)
    assert np.all(np.abs(oracle.vector_map("atanh", t) - a) < 1e-3 * np.maximum(1, np.abs(a) ** 4))     # test_atanh_tanh_inverse (tanh saturates)
    assert np.array_equal(oracle.vector_map("atanh", -t), -oracle.vector_map("atanh", t))


@CASES
@given(a=vec(-1.0, 1.0))
def test_inverse_trig(oracle, a):
    asin, acos = oracle.vector_map("asin", a), oracle.vector_map("acos", a)
    assert np.all(np.abs(np.sin(asin.astype(np.float64)) - a) < 1e-5)           # test_asin_sin_inverse
    assert np.all(np.abs(np.cos(acos.astype(np.float64)) - a) < 1e-5)           # test_acos_cos_inverse
    assert np.all(np.abs(asin + acos - f32(np.pi / 2)) < 1e-5)                  # test_acos_symmetry family
    assert np.array_equal(oracle.vector_map("asin", -a), -asin)                 # test_asin_odd_function
    assert np.all((asin >= -np.pi / 2 - 1e-6) & (asin <= np.pi / 2 + 1e-6)) and np.all((acos >= 0) & (acos <= np.pi + 1e-6))


@CASES
@given(a=vec(-100.0, 100.0))
def test_atan_trunc_fract_signum_neg(oracle, a):
    at = oracle.vector_map("atan", a)
    assert np.all(np.abs(at) < np.pi / 2) and np.array_equal(oracle.vector_map("atan", -a), -at)        # range, odd
    tr, fr = oracle.vector_map("trunc", a), oracle.vector_map("fract", a)
    assert np.array_equal(oracle.vector_map("trunc", tr), tr)                   # test_trunc_idempotence
    assert np.all(np.abs(tr) <= np.abs(a))                                      # test_trunc_toward_zero
    assert np.array_equal(tr + fr, a) and np.all(np.abs(fr) < 1)                # test_fract_decomposition / _magnitude
    sg = oracle.vector_map("signum", a)
    assert np.all(np.abs(sg) == 1)                                              # test_signum_range (no NaN here)
    big = np.abs(a) > 1e-10
    assert np.all(np.abs(sg[big] * np.abs(a[big]) - a[big]) < 1e-5)            # test_signum_abs_identity
    ng = oracle.vector_map("neg", a)
    assert np.array_equal(oracle.vector_map("neg", ng), a) and np.array_equal(np.abs(ng), np.abs(a))   # double negation, magnitude


@CASES
@given(data=st.data())
def test_binary_maps(oracle, data):
    n = data.draw(st.integers(1, 100))
    a = data.draw(hnp.arrays(f32, n, elements=st.floats(-100, 100, width=32)))
    b = data.draw(hnp.arrays(f32, n, elements=st.floats(-100, 100, width=32)))
    mn, mx = oracle.vector_map("minimum", a, b), oracle.vector_map("maximum", a, b)
    assert np.array_equal(mn, oracle.vector_map("minimum", b, a)) and np.array_equal(mx, oracle.vector_map("maximum", b, a))   # commutative
    assert np.array_equal(oracle.vector_map("minimum", a, a), a) and np.array_equal(oracle.vector_map("maximum", a, a), a)     # idempotent
    assert np.all(mn <= mx) and np.array_equal(mn + mx, a + b)
    cs = oracle.vector_map("copysign", a, b)
    assert np.array_equal(np.abs(cs), np.abs(a)) and np.array_equal(np.signbit(cs), np.signbit(b))   # magnitude kept, sign copied


@CASES
@given(a=vec(1.0, 10.0, max_len=50), n=st.floats(1, 3, width=32), m=st.floats(1, 3, width=32))
def test_pow_power_laws(oracle, a, n, m):                                       # test_pow_power_laws
    nm = oracle.vector_map("pow", a, p0=float(f32(n) * f32(m)))
    twice = oracle.vector_map("pow", oracle.vector_map("pow", a, p0=float(n)), p0=float(m))
    assert np.all(np.abs(nm - twice) < np.where(np.abs(nm) > 1, np.abs(nm) * 1e-3, 1e-3))
    assert np.array_equal(oracle.vector_map("pow", a, p0=1.0), a) and np.all(oracle.vector_map("pow", a, p0=0.0) == 1)   # test_pow_special_cases


@CASES
@given(a=vec(-100.0, 100.0), slope=st.floats(0.0, 0.984375, width=32))
def test_activation_shapes(oracle, a, slope):
    lr = oracle.vector_map("leaky_relu", a, p0=float(slope))
    assert np.array_equal(lr[a > 0], a[a > 0]) and np.all(lr[a <= 0] <= 0)
    hs = oracle.vector_map("hardswish", a)
    assert np.array_equal(hs[a >= 3], a[a >= 3]) and np.all(hs[a <= -3] == 0) and np.all(hs >= -0.375 - 1e-6)
    mi = oracle.vector_map("mish", a)
    assert np.all(mi >= -0.31) and np.array_equal(mi[a > 20], a[a > 20]) and np.all(mi[a < -20] == 0)
    se = oracle.vector_map("selu", a)
    assert np.all(se[a > 0] > 0) and np.all(se[a <= 0] <= 0) and np.all(se >= -1.7581 - 1e-4)
    el = oracle.vector_map("elu", a, p0=1.0)
    assert np.array_equal(el[a > 0], a[a > 0]) and np.all((el[a <= 0] >= -1.0) & (el[a <= 0] <= 0))
