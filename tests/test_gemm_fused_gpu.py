"""The fused-split tcgen05 GEMM (csrc/gemm_tc.cu `gemm_tf32x3_fused_pair_kernel`: raw operands through TMA, lo halves
computed in shared memory, B consumed MN-major — no pre-pass) against the f64 truth on the matmul contract
(|C - truth| <= 1e-5 * sum_k |a_ik||b_kj|, Matrix::matmul src/matrix.rs:285), against the pre-pass kernel, and on the
reference's behavioural tests (tests/wasm_optimization_tests.rs:117-230: prime dims, NaN / Inf propagation, bit-identical
reruns) with the kernel FORCED for every K (TRN_GEMM_FUSED is read once per process, so the forced runs are subprocesses).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
f32 = np.float32
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

# (batch, m, k, n): K tails (k % 16 != 0), n % 32 != 0, m tails inside a 256-row pair tile, one and many k-chunks
SHAPES = [(1, 256, 128, 256), (1, 129, 4, 4), (1, 300, 36, 260), (1, 257, 200, 516), (3, 384, 72, 132), (1, 1024, 1024, 1024),
          (2, 130, 2052, 300), (1, 2048, 128, 2048), (4, 512, 64, 512), (1, 131, 16, 100),
          # K <= 128 with several n-tiles: the A-stationary form (panel changes inside a pair's tile range, ragged edges)
          (2, 520, 128, 1100), (1, 256, 4, 516), (5, 300, 100, 772), (37, 256, 32, 512)]


def _worker():
    import trueno_b200 as trn
    trn.check(trn.lib.trn_cuda_init(0))
    trn.set_gemm_engine(trn.ENGINE_TC_3XTF32)
    out = {}
    worst = 0.0
    for (batch, m, k, n) in SHAPES:
        rng = np.random.default_rng(batch * 1000003 + m * 7 + k * 3 + n)
        for signed in (False, True):
            A = rng.uniform(-1 if signed else 0, 1, (batch, m, k)).astype(f32)
            B = rng.uniform(-1 if signed else 0, 1, (batch, k, n)).astype(f32)
            if batch == 1:
                C = trn.Matrix.from_vec(m, k, A[0]).matmul(trn.Matrix.from_vec(k, n, B[0])).to_numpy()[None]
            else:
                C = trn.Matrix.batched_matmul(A.ravel(), B.ravel(), batch, m, k, n).reshape(batch, m, n)
            truth = A.astype(np.float64) @ B.astype(np.float64)
            scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64)
            ratio = float(np.max(np.abs(C - truth) / np.maximum(scale, 1e-30)))
            worst = max(worst, ratio)
            assert ratio <= 1e-5, ((batch, m, k, n), signed, ratio)
            out[f"{batch}x{m}x{k}x{n}{'s' if signed else 'u'}"] = C
    # NaN / Inf: the splitter warps raise the device flag, the gated SIMT kernel recomputes with IEEE semantics
    n = 160
    A = np.ones((n, n), f32); A[2, 3] = np.nan
    C = trn.Matrix.from_vec(n, n, A).matmul(trn.Matrix.from_vec(n, n, np.ones((n, n), f32))).to_numpy()
    assert np.isnan(C[2]).all() and np.isfinite(np.delete(C, 2, 0)).all()
    A = np.ones((n, n), f32); A[1, 1] = np.inf
    C = trn.Matrix.from_vec(n, n, A).matmul(trn.Matrix.from_vec(n, n, np.ones((n, n), f32))).to_numpy()
    assert np.isposinf(C[1]).all() and np.isfinite(np.delete(C, 1, 0)).all()
    B = np.ones((n, n), f32); B[5, 7] = -np.inf
    C = trn.Matrix.from_vec(n, n, np.ones((n, n), f32)).matmul(trn.Matrix.from_vec(n, n, B)).to_numpy()
    assert np.isneginf(C[:, 7]).all() and np.isfinite(np.delete(C, 7, 1)).all()
    big = np.full((n, n), np.finfo(f32).max, f32)
    C = trn.Matrix.from_vec(n, n, big).matmul(trn.Matrix.from_vec(n, n, np.full((n, n), 2, f32))).to_numpy()
    assert (np.isinf(C) | np.isnan(C)).all()
    # exact small-integer products (benches/matrix_ops.rs:23-25 data): every term and sum is exact in f32
    i = np.arange(512 * 512)
    A = (i % 100).astype(f32).reshape(512, 512)
    B = ((i * 2) % 100).astype(f32).reshape(512, 512)
    C = trn.Matrix.from_vec(512, 512, A).matmul(trn.Matrix.from_vec(512, 512, B)).to_numpy()
    assert np.array_equal(C.astype(np.float64), A.astype(np.float64) @ B.astype(np.float64))
    # bit-identical reruns
    A = ((np.arange(256 * 128) % 97).astype(f32) * f32(0.01)).reshape(256, 128)
    B = ((np.arange(128 * 256) % 83).astype(f32) * f32(0.01)).reshape(128, 256)
    Am, Bm = trn.Matrix.from_vec(256, 128, A), trn.Matrix.from_vec(128, 256, B)
    first = Am.matmul(Bm).as_slice().tobytes()
    for _ in range(50):
        assert Am.matmul(Bm).as_slice().tobytes() == first
    # every head of a batched product is bit-identical to the stand-alone product
    rng = np.random.default_rng(5)
    A = rng.standard_normal((3, 256, 64)).astype(f32)
    B = rng.standard_normal((3, 64, 384)).astype(f32)
    Cb = trn.Matrix.batched_matmul(A.ravel(), B.ravel(), 3, 256, 64, 384).reshape(3, 256, 384)
    for h in range(3):
        assert np.array_equal(Cb[h], trn.Matrix.from_vec(256, 64, A[h]).matmul(trn.Matrix.from_vec(64, 384, B[h])).to_numpy())
    np.savez(sys.argv[2], worst=np.float64(worst), **out)
    print(f"fused={os.environ.get('TRN_GEMM_FUSED')} worst error / (sum|a||b|) = {worst:.3e}")


def _run(mode: str, path: str):
    env = dict(os.environ, TRN_GEMM_FUSED=mode, PYTHONPATH=ROOT)
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "worker", path], env=env, cwd=ROOT,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-4000:]
    return np.load(path)


def test_fused_kernel_forced_for_every_k_and_prepass_kernel_agree(tmp_path):
    fused = _run("1", str(tmp_path / "fused.npz"))
    prepass = _run("0", str(tmp_path / "prepass.npz"))
    assert float(fused["worst"]) <= 1e-5 and float(prepass["worst"]) <= 1e-5
    for (batch, m, k, n) in SHAPES:
        for tag in "us":
            key = f"{batch}x{m}x{k}x{n}{tag}"
            a, b = fused[key].astype(np.float64), prepass[key].astype(np.float64)
            # two f32-accurate evaluations of the same product: they differ by rounding only
            assert np.max(np.abs(a - b)) <= 4e-6 * max(1.0, float(np.max(np.abs(b)))), key


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "worker":
    sys.path.insert(0, ROOT)
    _worker()
