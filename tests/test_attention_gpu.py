"""Fused attention (trueno-gpu/src/kernels/attention.rs:27-125) on the CUDA path against
(a) the oracle composition of the reference's own CPU operators (transpose + matmul -> scale -> softmax -> matmul)
and (b) the f64 truth.  Stated tolerance, per output element (h, i, c), with P the exact softmax:

    |out - truth| <= (1e-5 + 2e-5 * kappa) * sum_j P_ij |V_jc|,   kappa = max_ij scale * sum_c |Q_ic| |K_jc|

i.e. the matmul contract (1e-5 * sum|terms|) for the P V product, plus the score perturbation the same contract
allows on scale * Q K^T (|dx| <= 1e-5 * kappa) passing through exp twice (numerator and denominator).  The oracle
is held to the same bound; CUDA vs oracle is therefore within twice it."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
f32 = np.float32


def truth64(q, k, v, heads, seq, d, scale, causal):
    q = q.reshape(heads, seq, d).astype(np.float64); k = k.reshape(heads, seq, d).astype(np.float64)
    v = v.reshape(heads, seq, d).astype(np.float64)
    s = np.einsum("hid,hjd->hij", q, k) * float(scale)
    if causal:
        s = np.where(np.arange(seq)[None, None, :] > np.arange(seq)[None, :, None], -np.inf, s)
    s -= s.max(axis=-1, keepdims=True)
    p = np.exp(s)
    p /= p.sum(axis=-1, keepdims=True)
    out = np.einsum("hij,hjd->hid", p, v)
    bound = np.einsum("hij,hjd->hid", p, np.abs(v))
    kappa = float(scale) * np.einsum("hid,hjd->hij", np.abs(q), np.abs(k)).max()
    return out.ravel(), bound.ravel(), kappa


def make(heads, seq, d, seed, spread=1.0):
    rng = np.random.default_rng(seed)
    n = heads * seq * d
    return ((rng.standard_normal(n) * spread).astype(f32), (rng.standard_normal(n) * spread).astype(f32),
            rng.standard_normal(n).astype(f32))


SHAPES = [(2, 4, 8), (1, 1, 1), (3, 128, 64), (2, 200, 128), (1, 384, 32), (2, 130, 16), (1, 77, 8), (1, 256, 100),
          (2, 129, 5), (1, 512, 128), (1, 40, 160), (1, 33, 300)]


@pytest.mark.parametrize("causal", [False, True], ids=["full", "causal"])
@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_attention_vs_oracle_and_truth(trn, oracle, shape, causal):
    heads, seq, d = shape
    q, k, v = make(heads, seq, d, seed=seq * 131 + d)
    scale = f32(1.0) / np.sqrt(f32(d))
    got = trn.attention(q, k, v, heads, seq, d, causal=causal)
    want, bound, kappa = truth64(q, k, v, heads, seq, d, scale, causal)
    tol = (1e-5 + 2e-5 * kappa) * bound + 1e-30
    assert np.all(np.isfinite(got))
    assert np.all(np.abs(got - want) <= tol), float(np.max(np.abs(got - want) / tol))
    ref = oracle.attention(q, k, v, heads, seq, d, causal=causal)
    assert np.all(np.abs(ref - want) <= tol)            # the oracle meets the same contract
    assert np.all(np.abs(got - ref) <= 2 * tol)


@pytest.mark.parametrize("engine", ["ENGINE_SIMT", "ENGINE_TC_3XTF32"])
@pytest.mark.parametrize("shape", [(2, 4, 8), (2, 200, 128), (1, 300, 64), (1, 129, 24)], ids=lambda s: "x".join(map(str, s)))
def test_attention_forced_engines(trn, shape, engine):
    heads, seq, d = shape
    q, k, v = make(heads, seq, d, seed=7, spread=2.0)
    trn.set_gemm_engine(getattr(trn, engine))
    try:
        for causal in (False, True):
            got = trn.attention(q, k, v, heads, seq, d, scale=0.2, causal=causal)
            want, bound, kappa = truth64(q, k, v, heads, seq, d, 0.2, causal)
            tol = (1e-5 + 2e-5 * kappa) * bound + 1e-30
            assert np.all(np.abs(got - want) <= tol), float(np.max(np.abs(got - want) / tol))
    finally:
        trn.set_gemm_engine(trn.ENGINE_AUTO)


def test_attention_uniform_scores_are_means(trn):
    """q = 0 -> every key weighs the same: out = mean of V (prefix means when causal)."""
    heads, seq, d = 2, 300, 64
    rng = np.random.default_rng(1)
    v = rng.standard_normal(heads * seq * d).astype(f32)
    k = rng.standard_normal(heads * seq * d).astype(f32)
    q = np.zeros(heads * seq * d, f32)
    v3 = v.reshape(heads, seq, d).astype(np.float64)
    got = trn.attention(q, k, v, heads, seq, d).reshape(heads, seq, d)
    mean = v3.mean(axis=1, keepdims=True)
    assert np.max(np.abs(got - mean)) <= 2e-6 * np.abs(v3).mean(axis=1).max() * 4
    got_c = trn.attention(q, k, v, heads, seq, d, causal=True).reshape(heads, seq, d)
    prefix = np.cumsum(v3, axis=1) / np.arange(1, seq + 1)[None, :, None]
    assert np.max(np.abs(got_c - prefix)) <= 1e-5


def test_attention_one_hot_selects_value_row(trn):
    """A key that dominates by 40 in the exponent selects its value row exactly (the other weights underflow to
    less than half an ulp of the result)."""
    heads, seq, d = 1, 256, 128
    rng = np.random.default_rng(2)
    v = rng.standard_normal(seq * d).astype(f32)
    target = np.arange(seq)[::-1].copy()      # query i attends to key seq-1-i
    # key j is coded by two one-hot coordinates (j % 64, 64 + j // 64); the matching key scores 80, others <= 40
    q = np.zeros((seq, d), f32); k = np.zeros((seq, d), f32)
    for i in range(seq):
        t = target[i]
        q[i, t % 64] = 40.0
        q[i, 64 + t // 64] = 40.0
    for j in range(seq):
        k[j, j % 64] = 1.0
        k[j, 64 + j // 64] = 1.0
    got = trn.attention(q.ravel(), k.ravel(), v, heads, seq, d, scale=1.0).reshape(seq, d)
    want = v.reshape(seq, d)[target]
    assert np.max(np.abs(got - want)) <= 1e-6 * np.abs(want).max()


def test_attention_is_deterministic(trn):
    heads, seq, d = 4, 333, 128
    q, k, v = make(heads, seq, d, seed=3)
    a = trn.attention(q, k, v, heads, seq, d, causal=True)
    b = trn.attention(q, k, v, heads, seq, d, causal=True)
    assert np.array_equal(a, b)


def test_attention_non_finite_inputs_take_the_ieee_path(trn):
    """Inf/NaN inputs: hi*lo would manufacture NaNs, so the tensor kernel hands over to the SIMT kernel (device
    flag).  The automatic path must then equal the forced SIMT engine bit for bit, NaNs in the same places."""
    heads, seq, d = 2, 160, 64
    q, k, v = make(heads, seq, d, seed=4)
    v[5 * d + 3] = np.inf
    q[(seq + 17) * d + 1] = np.nan
    auto = trn.attention(q, k, v, heads, seq, d)
    trn.set_gemm_engine(trn.ENGINE_SIMT)
    try:
        simt = trn.attention(q, k, v, heads, seq, d)
    finally:
        trn.set_gemm_engine(trn.ENGINE_AUTO)
    assert np.array_equal(np.isnan(auto), np.isnan(simt))
    ok = ~np.isnan(auto)
    assert np.array_equal(auto[ok], simt[ok])
    assert np.isnan(auto.reshape(heads, seq, d)[1, 17]).all()       # the NaN query row
    assert np.isfinite(auto.reshape(heads, seq, d)[1, 18]).all()    # other rows of that head are clean


def test_attention_errors(trn):
    z = np.zeros(2 * 4 * 8, f32)
    with pytest.raises(trn.TruenoError) as e:
        trn.attention(np.zeros(50, f32), z, z, 2, 4, 8)
    assert "Q data size mismatch: expected 64 (2×4×8), got 50" in str(e.value)
    with pytest.raises(trn.TruenoError) as e:
        trn.attention(z, z, np.zeros(3, f32), 2, 4, 8)
    assert "V data size mismatch" in str(e.value)
    big = np.zeros(1 * 2 * 2048, f32)
    with pytest.raises(trn.TruenoError) as e:
        trn.attention(big, big, big, 1, 2, 2048)
    assert "head_dim 2048 exceeds" in str(e.value)
    assert trn.attention(np.zeros(0, f32), np.zeros(0, f32), np.zeros(0, f32), 0, 4, 8).size == 0


def test_attention_full_size_rows_sum_to_one(trn):
    """BASELINE config 3's head shape (seq 2048, head_dim 128): with V = 1 every output is the row sum of the
    softmax = 1; with V = key index the output is the softmax-weighted mean index, inside [0, seq)."""
    heads, seq, d = 8, 2048, 128
    rng = np.random.default_rng(5)
    q = rng.standard_normal(heads * seq * d).astype(f32)
    k = rng.standard_normal(heads * seq * d).astype(f32)
    v = np.ones(heads * seq * d, f32)
    for causal in (False, True):
        out = trn.attention(q, k, v, heads, seq, d, causal=causal)
        assert np.max(np.abs(out - 1.0)) <= 4e-6


@pytest.mark.parametrize("mode", ["0", "2"], ids=["single-cta", "forced-pairs"])
def test_attention_kernel_variants_in_subprocess(trn, mode):
    """The launcher picks the CTA-pair kernel or the single-CTA kernel by shape; TRN_ATT_PAIR (read once per process)
    forces either, so both run the multi-tile, ragged and causal shapes here, each in its own process."""
    import os
    import subprocess
    import sys
    code = r'''
import sys, numpy as np
sys.path.insert(0, "tests"); sys.path.insert(0, ".")
import trueno_b200 as trn
from test_attention_gpu import make, truth64
for heads, seq, d in [(2, 200, 128), (1, 384, 32), (2, 129, 5), (1, 640, 64), (3, 300, 96)]:
    for causal in (False, True):
        q, k, v = make(heads, seq, d, seed=seq + d)
        scale = np.float32(1.0) / np.sqrt(np.float32(d))
        got = trn.attention(q, k, v, heads, seq, d, causal=causal)
        want, bound, kappa = truth64(q, k, v, heads, seq, d, scale, causal)
        tol = (1e-5 + 2e-5 * kappa) * bound + 1e-30
        assert np.all(np.abs(got - want) <= tol), (heads, seq, d, causal, float(np.max(np.abs(got - want) / tol)))
print("variants ok")
'''
    env = dict(os.environ, TRN_ATT_PAIR=mode)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "variants ok" in r.stdout, r.stdout + r.stderr
