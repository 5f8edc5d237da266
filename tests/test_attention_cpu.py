"""The attention oracle (composition of the reference's CPU operators, oracle/__init__.py::attention) against the
f64 truth and hand-checkable cases — CPU only."""
import numpy as np
import pytest

from test_attention_gpu import make, truth64

f32 = np.float32


@pytest.mark.parametrize("causal", [False, True])
@pytest.mark.parametrize("shape", [(2, 4, 8), (1, 1, 1), (2, 67, 16), (1, 130, 64)], ids=lambda s: "x".join(map(str, s)))
def test_oracle_attention_vs_truth(oracle, shape, causal):
    heads, seq, d = shape
    q, k, v = make(heads, seq, d, seed=11)
    ref = oracle.attention(q, k, v, heads, seq, d, causal=causal)
    want, bound, kappa = truth64(q, k, v, heads, seq, d, f32(1.0) / np.sqrt(f32(d)), causal)
    assert np.all(np.abs(ref - want) <= (1e-5 + 2e-5 * kappa) * bound + 1e-30)


def test_oracle_attention_uniform_and_causal_prefix(oracle):
    heads, seq, d = 1, 5, 2
    v = np.arange(10, dtype=f32)
    z = np.zeros(10, f32)
    out = oracle.attention(z, z, v, heads, seq, d).reshape(seq, d)
    assert np.allclose(out, [[4, 5]] * 5, atol=1e-6)
    out = oracle.attention(z, z, v, heads, seq, d, causal=True).reshape(seq, d)
    assert np.allclose(out, [[0, 1], [1, 2], [2, 3], [3, 4], [4, 5]], atol=1e-6)
