"""Second, independent restatement of the AVX2 accumulation ORDER, against which the C oracle is held bit for bit.

The reference's own known-answer tests for Avx2Backend::dot / sum have at most 9 elements (src/backends/avx2.rs:1654-1717):
they run one 8-lane step plus the scalar tail and never reach the 4-accumulator 32-element loop (avx2.rs:170-187).  The
reference cannot be run here (Rust), so that loop is pinned the only other way there is: a model of avx2.rs:159-216 (dot)
and :225-252 (sum) written from the source in exact rational arithmetic — every f32 operation rounded once, FMA as ONE
rounding — on data whose result depends on the order (wide exponent range, cancellation).  Two restatements written
independently, in different languages, agreeing in every bit, is not reference output; it is what rules out a transcription
slip in oracle/trueno_oracle.c, and the header of the oracle says which of its functions are pinned how."""
from fractions import Fraction

import numpy as np
import pytest

f32 = np.float32


def rnd(q: Fraction) -> np.float32:
    """q rounded to the nearest f32, ties to even (exact: decided on rationals, not through a double)."""
    if q == 0:
        return f32(0.0)
    c = f32(float(q))                      # a candidate within one ulp (float(q) is correctly rounded to f64)
    best = c
    for cand in (np.nextafter(c, f32(-np.inf)), np.nextafter(c, f32(np.inf))):
        d_best, d_cand = abs(Fraction(float(best)) - q), abs(Fraction(float(cand)) - q)
        if d_cand < d_best or (d_cand == d_best and (int(cand.view(np.uint32)) & 1) == 0 and (int(best.view(np.uint32)) & 1) == 1):
            best = cand
    return best


def F(x) -> Fraction:
    return Fraction(float(x))


def fadd(a, b):
    return rnd(F(a) + F(b))


def fmul(a, b):
    return rnd(F(a) * F(b))


def fma(a, b, c):
    return rnd(F(a) * F(b) + F(c))


def hsum8(acc):
    # avx2.rs:203-210: halves, movehl, shuffle
    sh = [fadd(acc[j], acc[j + 4]) for j in range(4)]
    t0, t1 = fadd(sh[0], sh[2]), fadd(sh[1], sh[3])
    return fadd(t0, t1)


def model_dot(a, b):
    n, i = len(a), 0
    acc = [[f32(0)] * 8 for _ in range(4)]
    while i + 32 <= n:                                   # avx2.rs:170-187
        for u in range(4):
            acc[u] = [fma(a[i + 8 * u + l], b[i + 8 * u + l], acc[u][l]) for l in range(8)]
        i += 32
    while i + 8 <= n:                                    # :190-195, on acc0
        acc[0] = [fma(a[i + l], b[i + l], acc[0][l]) for l in range(8)]
        i += 8
    a01 = [fadd(acc[0][l], acc[1][l]) for l in range(8)]  # :198-200
    a23 = [fadd(acc[2][l], acc[3][l]) for l in range(8)]
    res = hsum8([fadd(a01[l], a23[l]) for l in range(8)])
    tail = f32(0)
    for j in range(i, n):                                # :213: x * y rounded, then summed sequentially
        tail = fadd(tail, fmul(a[j], b[j]))
    return fadd(res, tail)


def model_sum(a):
    n, i = len(a), 0
    acc = [f32(0)] * 8
    while i + 8 <= n:                                    # avx2.rs:232-236: ONE accumulator
        acc = [fadd(acc[l], a[i + l]) for l in range(8)]
        i += 8
    res = hsum8(acc)
    tail = f32(0)
    for j in range(i, n):
        tail = fadd(tail, a[j])
    return fadd(res, tail)


def _data(n, seed):
    rng = np.random.default_rng(seed)
    # exponents over 2^-12 .. 2^12 and both signs: the f32 result depends on the association
    return (rng.standard_normal(n) * np.exp2(rng.integers(-12, 13, n))).astype(f32)


@pytest.mark.parametrize("n", [1, 7, 8, 9, 31, 32, 33, 40, 63, 64, 71, 100, 257, 1000])
def test_oracle_avx2_dot_and_sum_follow_the_reference_order_bit_for_bit(oracle, n):
    from oracle import AVX2
    for seed in (n, n + 1000):
        a, b = _data(n, seed), _data(n, seed + 1)
        got_dot, want_dot = oracle.dot(a, b, backend=AVX2), model_dot(a, b)
        assert got_dot.view(np.uint32) == want_dot.view(np.uint32), (n, seed, got_dot, want_dot)
        got_sum, want_sum = oracle.sum(a, backend=AVX2), model_sum(a)
        assert got_sum.view(np.uint32) == want_sum.view(np.uint32), (n, seed, got_sum, want_sum)
        # and the order matters on this data: a plain left-to-right f32 sum differs somewhere in the sweep
    a = _data(257, 5)
    seq = f32(0)
    for v in a:
        seq = fadd(seq, v)
    assert seq.view(np.uint32) != model_sum(a).view(np.uint32)


def test_norm_l2_is_sqrt_of_dot_with_itself(oracle):
    from oracle import AVX2
    a = _data(100, 9)                                    # avx2.rs:481-489
    assert oracle.norm_l2(a, backend=AVX2) == np.sqrt(model_dot(a, a))
