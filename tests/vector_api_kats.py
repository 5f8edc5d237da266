"""Known-answer tests the reference holds for the rest of Vector's element-wise / statistics API
(src/vector.rs tests, lines cited per case), written once and run against the oracle (CPU suite) and against the
CUDA path (GPU suite) through a small adapter: api.call(op, *vectors, **params) -> np.ndarray | float, raising an
exception with a `.variant` attribute for TruenoError values."""
import math

import numpy as np

f32 = np.float32
LAMBDA = f32(1.0507009873554804934193349852946)
ALPHA = f32(1.6732632423543772848170429916717)
PI = math.pi


def _close(got, want, tol):
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    assert got.shape == want.shape and np.all(np.abs(got - want) < tol), (got, want)


def _raises(fn, variant):
    try:
        fn()
    except Exception as e:  # noqa: BLE001
        assert getattr(e, "variant", None) == variant, (e, variant)
        return e
    raise AssertionError(f"expected {variant}")


def run(call):
    # ---- activations -------------------------------------------------------------------------------
    assert np.array_equal(call("leaky_relu", [-2, -1, 0, 1, 2], p=(0.01,)), np.array([-0.02, -0.01, 0, 1, 2], f32))  # :8170
    for slope, want in ((0.01, -0.1), (0.1, -1.0), (0.2, -2.0)):                                                    # :8179
        r = call("leaky_relu", [-10, 5], p=(slope,))
        assert abs(r[0] - want) < 1e-6 and r[1] == 5.0
    assert np.array_equal(call("leaky_relu", [-3, 2], p=(0.0,)), np.array([-0.0, 2], f32))                          # :8200 (relu)
    _raises(lambda: call("leaky_relu", [], p=(0.01,)), "EmptyVector")                                               # :8226
    for bad in (-0.1, 1.0, 1.5):                                                                                    # :8233
        _raises(lambda: call("leaky_relu", [1, 2, 3], p=(bad,)), "InvalidInput")
    r = call("elu", [-2, -1, 0, 1, 2], p=(1.0,))                                                                    # :8249
    assert abs(r[0] + 0.8647) < 1e-3 and abs(r[1] + 0.6321) < 1e-3 and r[2] == 0 and r[3] == 1 and r[4] == 2
    _close(call("elu", [-10, -20, -100], p=(1.0,)), [-1, -1, -1], 1e-3)                                             # :8285
    _raises(lambda: call("elu", [], p=(1.0,)), "EmptyVector")                                                       # :8315
    for bad in (0.0, -1.0):                                                                                         # :8322
        _raises(lambda: call("elu", [1, 2, 3], p=(bad,)), "InvalidInput")
    r = call("hardswish", [-4, -3, -1.5, 0, 1.5, 3, 4])                                                             # :8492
    assert r[0] == 0 and r[1] == 0 and abs(r[2] + 0.375) < 1e-5 and r[3] == 0 and abs(r[4] - 1.125) < 1e-5 and r[5] == 3 and r[6] == 4
    _close(call("hardswish", [-2, -1, 1, 2]), [-1 / 3, -1 / 3, 2 / 3, 5 / 3], 1e-5)                                 # :8547
    _raises(lambda: call("hardswish", []), "EmptyVector")                                                           # :8563
    r = call("mish", [-2, -1, 0, 1, 2])                                                                             # :8571
    assert r[0] < 0 and r[1] < 0 and abs(r[2]) < 1e-5 and r[3] > 0 and r[4] > 0
    r = call("mish", [-1.19])                                                                                       # :8620
    assert -0.4 < r[0] < -0.2
    _close(call("mish", [25.0, 50.0]), [25.0, 50.0], 1e-5)                                                          # :8596
    _close(call("mish", [-25.0, -50.0]), [0.0, 0.0], 1e-5)                                                          # :8608
    _raises(lambda: call("mish", []), "EmptyVector")                                                                # :8631
    r = call("selu", [-2, -1, 0, 1, 2])                                                                             # :8640
    assert abs(r[3] - LAMBDA) < 1e-5 and abs(r[4] - 2 * LAMBDA) < 1e-5 and abs(r[2]) < 1e-5
    assert abs(r[1] - LAMBDA * ALPHA * (np.exp(f32(-1)) - f32(1))) < 1e-5
    assert abs(call("selu", [-100.0])[0] + LAMBDA * ALPHA) < 1e-4                                                   # :8689
    _raises(lambda: call("selu", []), "EmptyVector")                                                                # :8721
    assert np.array_equal(call("clip", [-5, 0, 5, 10, 15], p=(0.0, 10.0)), np.array([0, 0, 5, 10, 10], f32))        # :7789
    assert np.array_equal(call("clip", [1, 2, 3], p=(2.0, 2.0)), np.array([2, 2, 2], f32))                          # :7834
    e = _raises(lambda: call("clip", [1, 2, 3], p=(10.0, 5.0)), "InvalidInput")                                     # :7825
    assert "min_val (10) must be <= max_val (5)" in str(e)
    # ---- plain maps ----------------------------------------------------------------------------------
    r = call("signum", [5.0, -3.0, 0.0, -0.0])                                                                      # test_signum_basic
    assert np.array_equal(r, np.array([1, -1, 1, -1], f32))
    assert np.isnan(call("signum", [np.nan])[0])
    assert np.array_equal(call("trunc", [3.2, 3.7, -2.3, -2.8, 5.0]), np.array([3, 3, -2, -2, 5], f32))             # test_trunc_basic
    _close(call("fract", [3.7, -2.3, 5.0]), [0.7, -0.3, 0.0], 1e-5)                                                 # test_fract_basic
    assert np.array_equal(call("copysign", [5, 3, 2, 4], [-1, 1, -1, 1]), np.array([-5, 3, -2, 4], f32))            # test_copysign_basic
    assert np.array_equal(call("minimum", [1, 5, 3, 2], [2, 3, 4, 1]), np.array([1, 3, 3, 1], f32))                 # test_minimum_basic
    assert np.array_equal(call("maximum", [1, 5, 3, 2], [2, 3, 4, 1]), np.array([2, 5, 4, 2], f32))                 # test_maximum_basic
    r = call("minimum", [np.nan, 5.0, np.nan], [3.0, np.nan, np.nan])                                               # test_minimum_nan
    assert r[0] == 3 and r[1] == 5 and np.isnan(r[2])
    _raises(lambda: call("minimum", [1, 2], [1, 2, 3]), "SizeMismatch")
    _raises(lambda: call("copysign", [1, 2], [1]), "SizeMismatch")
    assert np.array_equal(call("neg", [1, -2, 3, -4]), np.array([-1, 2, -3, 4], f32))                               # test_neg_basic
    assert np.array_equal(call("pow", [2, 3, 4, 5], p=(2.0,)), np.array([4, 9, 16, 25], f32))                       # test_pow_basic
    assert np.array_equal(call("pow", [4, 9, 16], p=(0.5,)), np.array([2, 3, 4], f32))                              # test_pow_fractional
    assert np.array_equal(call("pow", [2, 4, 10], p=(-1.0,)), np.array([0.5, 0.25, 0.1], f32))                      # test_pow_negative_exponent
    assert np.array_equal(call("pow", [2, 3, 4], p=(0.0,)), np.array([1, 1, 1], f32))                               # test_pow_zero_exponent
    assert np.array_equal(call("pow", [2, 3, 4], p=(1.0,)), np.array([2, 3, 4], f32))                               # test_pow_one_exponent
    _close(call("pow", [2, 3], p=(3.0,)), [8, 27], 1e-5)                                                            # test_pow_cube
    assert np.array_equal(call("copysign", [5, 5], [np.inf, -np.inf]), np.array([5, -5], f32))                      # test_copysign_infinity
    assert np.array_equal(call("copysign", [3, 3], [0.0, -0.0]), np.array([3, -3], f32))                            # test_copysign_zero
    assert np.array_equal(call("minimum", [np.inf, 5, -np.inf], [3, np.inf, -10]), np.array([3, 5, -np.inf], f32))  # test_minimum_infinity
    r = call("maximum", [np.nan, 5.0, np.nan], [3.0, np.nan, np.nan])                                               # test_maximum_nan
    assert r[0] == 3 and r[1] == 5 and np.isnan(r[2])
    r = call("neg", [0.0, -0.0])                                                                                    # test_neg_zero
    assert np.signbit(r[0]) and not np.signbit(r[1]) and r[0] == 0 and r[1] == 0
    r = call("neg", [np.nan, 5.0])                                                                                  # test_neg_nan
    assert np.isnan(r[0]) and r[1] == -5
    assert np.array_equal(call("neg", [np.inf, -np.inf]), np.array([-np.inf, np.inf], f32))                         # test_neg_infinity
    _close(call("fract", [-1.2, -2.5, -3.9]), [-0.2, -0.5, -0.9], 1e-5)                                             # test_fract_negative
    assert np.array_equal(call("trunc", [2.7, -2.7, 5.3, -5.3]), np.array([2, -2, 5, -5], f32))                     # test_trunc_toward_zero
    assert abs(call("acosh", [1.0])[0]) < 1e-5                                                                      # test_acosh_one
    assert np.isnan(call("acosh", [0.5])[0]) and np.isnan(call("asin", [1.5])[0])                                   # outside the domain -> NaN
    _close(call("sinh", [0, 1, -1]), [0, math.sinh(1), -math.sinh(1)], 1e-5)
    _close(call("cosh", [0, 1, -1]), [1, math.cosh(1), math.cosh(1)], 1e-5)
    _close(call("asin", [0, 1, -1, 0.5]), [0, PI / 2, -PI / 2, PI / 6], 1e-5)
    _close(call("acos", [0, 1, -1, 0.5]), [PI / 2, 0, PI, PI / 3], 1e-5)
    _close(call("atan", [0, 1, -1, 1.732]), [0, PI / 4, -PI / 4, PI / 3], 1e-3)
    _close(call("asinh", [0, 1, -1]), [0, math.asinh(1), -math.asinh(1)], 1e-5)
    _close(call("acosh", [1, 2, 3]), [0, math.acosh(2), math.acosh(3)], 1e-5)
    _close(call("atanh", [0, 0.5, -0.5]), [0, math.atanh(0.5), -math.atanh(0.5)], 1e-5)
    for op in ("neg", "signum", "trunc", "fract", "sinh", "atanh", "pow"):                                         # empty in -> empty out
        assert call(op, [], p=((2.0,) if op == "pow" else ())).size == 0
    # ---- statistics ----------------------------------------------------------------------------------
    assert call("sum_of_squares", [3, 4]) == 25.0 and call("sum_of_squares", []) == 0.0                             # :7219, :7227
    assert abs(call("covariance", [1, 2, 3], [2, 4, 6]) - 4 / 3) < 1e-5                                             # :7386
    assert abs(call("covariance", [1, 2, 3], [3, 2, 1]) + 2 / 3) < 1e-5                                             # :7396
    _raises(lambda: call("covariance", [1, 2], [1, 2, 3]), "SizeMismatch")                                          # :7423
    _raises(lambda: call("covariance", [], []), "EmptyVector")                                                      # :7437
    assert abs(call("correlation", [1, 2, 3, 4], [2, 4, 6, 8]) - 1) < 1e-5                                          # :7449
    assert abs(call("correlation", [1, 2, 3, 4], [4, 3, 2, 1]) + 1) < 1e-5                                          # :7458
    _raises(lambda: call("correlation", [5, 5, 5], [1, 2, 3]), "DivisionByZero")                                    # :7484
    z = call("zscore", [1, 2, 3, 4, 5])                                                                             # :7511
    assert abs(z.mean()) < 1e-5 and abs(z.std() - 1) < 1e-5
    _raises(lambda: call("zscore", [3, 3, 3, 3]), "DivisionByZero")                                                 # :7547
    _raises(lambda: call("zscore", []), "EmptyVector")                                                              # :7555
    _close(call("minmax_normalize", [1, 2, 3, 4, 5]), [0, 0.25, 0.5, 0.75, 1], 1e-5)                                # :7580
    _close(call("minmax_normalize", [-10, 0, 10]), [0, 0.5, 1], 1e-5)                                               # :7600
    _raises(lambda: call("minmax_normalize", [7, 7, 7]), "DivisionByZero")                                          # :7624
    _raises(lambda: call("minmax_normalize", []), "EmptyVector")                                                    # :7632
    y = call("layer_norm_simple", [1, 2, 3, 4], p=(1e-5,))                                                          # :7735
    assert abs(y.mean()) < 1e-5 and abs(y.std() - 1) < 1e-2
    _raises(lambda: call("layer_norm_simple", [], p=(1e-5,)), "EmptyVector")
