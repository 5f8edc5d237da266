"""Writes the committed golden fixtures of tests/golden/.

The reference (paiml/trueno) is Rust and cannot run in the build image, so there are no reference-generated
outputs.  What CAN be pinned, and is:
  reference_kats.json   every exact-value known-answer test of the reference's own test-suite for this path
                        (inputs, expected values, tolerance, reference file:line), serialised from tests/kats.py;
  seeded_fixtures.npz   the ORACLE's outputs (scalar-backend restatement, oracle/trueno_oracle.c) on the
                        reference's seeded fixtures — xorshift64 softmax vectors (tests/pixel_fkr.rs:132-153,
                        seeds 22222 / 34567), the `i % 100` matmul fixture at 256 (src/matrix.rs:2540-2600), the
                        sin/cos dot and norm fixtures (tests/smoke_e2e.rs:54-91) and a splitmix slice of every
                        BASELINE config — so a later change to the oracle or the kernels is caught against a
                        frozen file, not against a moving oracle.
  vector_api_fixtures.npz  the oracle's outputs for the widened Vector / Matrix API (22 maps, statistics,
                        embedding_lookup) on the same xorshift fixtures, frozen the same way.
Run from the repository root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import kats  # noqa: E402
import oracle  # noqa: E402
from oracle import SCALAR  # noqa: E402

f32 = np.float32


def tolist(x):
    if isinstance(x, np.ndarray):
        return [float(v) for v in x.reshape(-1)]
    if isinstance(x, (list, tuple)):
        return [tolist(v) for v in x]
    if isinstance(x, (np.floating, np.integer)):
        return float(x)
    return x


def main():
    doc = {
        "source": "exact-value KATs of paiml/trueno's own tests for the hot path (SURVEY.md 8c), via tests/kats.py",
        "reductions": [dict(name=n, op=op, args=tolist(list(a)), expected=e, tol=t, cite=c) for n, op, a, e, t, c in kats.REDUCTION_KATS],
        "arg": [dict(name=n, op=op, v=tolist(v), expected=e, cite=c) for n, op, v, e, c in kats.ARG_KATS],
        "matmul": [dict(name=n, a=tolist(a), a_shape=list(a.shape), b=tolist(b), b_shape=list(b.shape), expected=tolist(e), tol=t, cite=c)
                   for n, a, b, e, t, c in kats.MATMUL_KATS],
        "batched": {k: tolist(v) for k, v in kats.BATCHED_KAT.items()},
        "batched4d": {k: tolist(v) for k, v in kats.BATCHED4D_KAT.items()},
        "matvec": {k: tolist(v) for k, v in kats.MATVEC_KAT.items()},
        "map_ext": [dict(op=o, inputs=tolist(i), params=list(p), expected=tolist(e), tol=t, cite=c) for o, i, p, e, t, c in kats.MAP_EXT_KATS],
        "reduce_ext": [dict(op=o, v=tolist(v), expected=e, tol=t, cite=c) for o, v, e, t, c in kats.REDUCE_EXT_KATS],
    }
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump(doc, f, indent=1)

    orc = oracle.get()
    out = {}
    for seed in (22222, 34567):
        x = kats.SimpleRng(seed).gen_vec(2048)
        out[f"xorshift_{seed}_x"] = x
        out[f"xorshift_{seed}_softmax"] = orc.softmax(x, backend=SCALAR)
        out[f"xorshift_{seed}_log_softmax"] = orc.log_softmax(x, backend=SCALAR)
        out[f"xorshift_{seed}_sigmoid"] = orc.sigmoid(x, backend=SCALAR)
        out[f"xorshift_{seed}_gelu"] = orc.gelu(x, backend=SCALAR)
    A, B = kats.fixture_mod(256, 256, 256, 100, 10.0, 7, 100, 10.0)
    out["matmul_mod100_256"] = orc.matmul(A, A.shape, B, B.shape).reshape(256, 256)
    i = np.arange(10000, dtype=f32)
    a, b = np.sin(i).astype(f32), np.cos(i).astype(f32)
    out["dot_sincos_10000"] = np.array([orc.dot(a, b, backend=SCALAR)], f32)
    out["norm_sin_10000"] = np.array([orc.norm_l2(np.sin(f32(0.01) * i).astype(f32), backend=SCALAR)], f32)
    s = kats.splitmix_u01(0x5EED0005, 0, 1 << 16) * f32(2) - f32(1)
    t = kats.splitmix_u01(0x5EED0006, 0, 1 << 16) * f32(2) - f32(1)
    out["splitmix_slice_sum_dot_norm"] = np.array([orc.sum(s, backend=SCALAR), orc.dot(s, t, backend=SCALAR), orc.norm_l2(s, backend=SCALAR)], f32)
    out["splitmix_slice_argmax_argmin"] = np.array([orc.argmax(s), orc.argmin(s)], np.int64)
    np.savez_compressed(os.path.join(HERE, "seeded_fixtures.npz"), **out)

    # ---- the widened Vector / Matrix API (a separate file: seeded_fixtures.npz stays byte-identical) ----
    api = {}
    x = kats.SimpleRng(22222).gen_vec(2048) * f32(4)        # xorshift fixture scaled to (-4, 4)
    y = kats.SimpleRng(34567).gen_vec(2048) * f32(4)
    api["x"], api["y"] = x, y
    for op in ("neg", "signum", "trunc", "fract", "hardswish", "mish", "selu", "sinh", "cosh", "atan", "asinh"):
        api[op] = orc.vector_map(op, x)
    u = (x / f32(4.5)).astype(f32)                           # inside (-1, 1)
    api["u"] = u
    for op in ("asin", "acos", "atanh"):
        api[op] = orc.vector_map(op, u)
    api["acosh"] = orc.vector_map("acosh", (np.abs(x) + f32(1)).astype(f32))
    api["leaky_relu_0.01"] = orc.vector_map("leaky_relu", x, p0=0.01)
    api["elu_1.5"] = orc.vector_map("elu", x, p0=1.5)
    api["pow_2"] = orc.vector_map("pow", x, p0=2.0)
    api["clip_-0.5_0.75"] = orc.vector_map("clip", x, p0=-0.5, p1=0.75)
    for op in ("minimum", "maximum", "copysign"):
        api[op] = orc.vector_map(op, x, y)
    api["minmax_normalize"] = orc.minmax_normalize(x)
    api["zscore"] = orc.zscore(x, backend=SCALAR)
    api["layer_norm_simple_1e-5"] = orc.layer_norm_simple(x, 1e-5, backend=SCALAR)
    api["stats"] = np.array([orc.dot(x, x, backend=SCALAR), orc.covariance(x, y, backend=SCALAR),
                             orc.correlation(x, y, backend=SCALAR)], f32)
    idx = (np.arange(97, dtype=np.uint64) * 37) % 64
    api["embedding_idx"] = idx
    api["embedding_64x32"] = orc.embedding_lookup(x, 64, 32, idx)
    np.savez_compressed(os.path.join(HERE, "vector_api_fixtures.npz"), **api)
    print("wrote", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
