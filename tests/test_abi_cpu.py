"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/trueno_cuda.h declares, validates arguments with the reference's exact error values
BEFORE touching the device, and never falls back to the CPU."""
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "trueno_cuda.h")).read()
    return sorted(set(re.findall(r"TRN_API\s+[\w\s\*]+?\b(trn_\w+)\s*\(", text)))


def test_header_symbols_are_exported(trn):
    syms = declared_symbols()
    assert len(syms) >= 50
    for s in syms:
        assert hasattr(trn.lib, s), f"{s} declared in include/trueno_cuda.h but not exported"
    assert set(syms) == set(trn.EXPORTED_SYMBOLS), set(syms) ^ set(trn.EXPORTED_SYMBOLS)


def test_no_oracle_or_cpu_fallback_in_product():
    # the product package must not import or link the oracle (prompt ③)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "trueno_b200")):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                src = open(os.path.join(dirpath, fn), errors="replace").read()
                assert "import oracle" not in src and "liboracle" not in src and "from oracle" not in src, fn


def test_validation_errors_match_reference(trn):
    V, M, E = trn.Vector, trn.Matrix, trn.TruenoError
    # src/vector.rs:589-594
    with pytest.raises(E) as e:
        V.from_slice([1, 2, 3]).dot(V.from_slice([1, 2]))
    assert e.value == E.SizeMismatch(3, 2) and str(e.value) == "Size mismatch: expected 3, got 2"
    with pytest.raises(E) as e:
        V.from_slice([1, 2, 3]).add(V.from_slice([1, 2, 3, 4]))
    assert e.value == E.SizeMismatch(3, 4)
    # src/vector.rs:4849-4857 — equality with InvalidInput("Empty vector")
    for op in ("max", "min", "argmax", "argmin"):
        with pytest.raises(E) as e:
            getattr(V.from_slice([]), op)()
        assert e.value == E.InvalidInput("Empty vector")
    # src/vector.rs:7908, 8152, 8413
    for op in ("softmax", "log_softmax", "sigmoid", "gelu"):
        with pytest.raises(E) as e:
            getattr(V.from_slice([]), op)()
        assert e.value == E.EmptyVector
    # src/matrix.rs:286-291
    with pytest.raises(E) as e:
        M.from_vec(2, 3, np.ones(6)).matmul(M.from_vec(2, 2, np.ones(4)))
    assert e.value.variant == "InvalidInput" and e.value.message == (
        "Matrix dimension mismatch for multiplication: 2×3 × 2×2 (inner dimensions 3 and 2 must match)")
    # src/matrix.rs:108-117
    with pytest.raises(E) as e:
        M.from_vec(2, 2, np.ones(3))
    assert e.value.message == "Data length 3 does not match matrix dimensions 2x2 (expected 4)"
    # src/matrix.rs:3912-3945, 4008-4043
    with pytest.raises(E) as e:
        M.batched_matmul(np.ones(10), np.ones(12), 2, 2, 3, 2)
    assert e.value.message == "A data size mismatch: expected 12 (2×2×3), got 10"
    with pytest.raises(E) as e:
        M.batched_matmul_4d(np.ones(8), np.ones(7), 1, 2, 2, 2, 2)
    assert e.value.message == "B data size mismatch: expected 8 (1×2×2×2), got 7"
    # src/matrix.rs:1658-1664
    with pytest.raises(E) as e:
        M.from_vec(2, 3, np.ones(6)).matvec(V.from_slice([1, 2]))
    assert e.value.message == "Vector length 2 does not match matrix columns 3 for matrix-vector multiplication"
    # row blocks of a sharded product: the mismatch text names the WHOLE matrix (src/matrix.rs:286-291), and a block
    # cannot be taller than it — both checked before any device work
    L = trn.lib
    with pytest.raises(E) as e:
        trn.check(L.trn_matmul_rowblock_f32_dev(None, 2, 8, 3, None, 2, 2, None, None))
    assert e.value.message == "Matrix dimension mismatch for multiplication: 8×3 × 2×2 (inner dimensions 3 and 2 must match)"
    with pytest.raises(E) as e:
        trn.check(L.trn_matmul_rowblock_f32_dev(None, 9, 8, 3, None, 3, 2, None, None))
    assert e.value.variant == "InvalidInput" and e.value.message == "row block of 9 rows exceeds the 8 rows of the matrix"
    with pytest.raises(E) as e:
        trn.check(L.trn_matmul_rowblock_prepared_f32_dev(None, 9, 8, 3, None, None, None))
    assert e.value.message == "row block of 9 rows exceeds the 8 rows of the matrix"


def test_compute_fails_loudly_without_gpu(trn):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-GPU contract is checked on CPU boxes")
    assert not trn.is_available()
    with pytest.raises(trn.TruenoError) as e:
        trn.Vector.from_slice([1, 2, 3]).sum()
    assert e.value.variant == "GpuError" and "no CPU fallback" in e.value.message
    with pytest.raises(trn.TruenoError) as e:
        trn.Matrix.identity(4).matmul(trn.Matrix.identity(4))
    assert e.value.variant == "GpuError"


def test_widened_api_validation_errors_match_reference(trn):
    """The error contract of the widened Vector / Matrix API is checked before any device work, so it holds on a box
    without a GPU too (values and message text of src/vector.rs / src/matrix.rs)."""
    V, M, E = trn.Vector, trn.Matrix, trn.TruenoError
    for op in ("hardswish", "mish", "selu", "zscore", "minmax_normalize"):       # src/vector.rs:2410, :2478, :2547, :1181, :1249
        with pytest.raises(E) as e:
            getattr(V.from_slice([]), op)()
        assert e.value == E.EmptyVector, op
    with pytest.raises(E) as e:
        V.from_slice([]).leaky_relu(0.01)                                         # :1981
    assert e.value == E.EmptyVector
    with pytest.raises(E) as e:
        V.from_slice([]).elu(1.0)                                                 # :2086
    assert e.value == E.EmptyVector
    with pytest.raises(E) as e:
        V.from_slice([]).layer_norm_simple(1e-5)                                  # :1387
    assert e.value == E.EmptyVector
    for bad, text in ((-0.1, "-0.1"), (1.0, "1"), (1.5, "1.5")):                  # :1986-1990
        with pytest.raises(E) as e:
            V.from_slice([1, 2, 3]).leaky_relu(bad)
        assert e.value == E.InvalidInput(f"negative_slope must be in [0.0, 1.0), got {text}")
    for bad, text in ((0.0, "0"), (-1.0, "-1")):                                  # :2091-2095
        with pytest.raises(E) as e:
            V.from_slice([1, 2, 3]).elu(bad)
        assert e.value == E.InvalidInput(f"alpha must be > 0, got {text}")
    with pytest.raises(E) as e:
        V.from_slice([1, 2, 3]).clip(10.0, 5.0)                                   # :1449-1454
    assert e.value == E.InvalidInput("min_val (10) must be <= max_val (5)")
    for op in ("minimum", "maximum", "copysign", "covariance", "correlation"):   # :4329, :4365, :4293, :1067, :1119
        with pytest.raises(E) as e:
            getattr(V.from_slice([1, 2]), op)(V.from_slice([1, 2, 3]))
        assert e.value == E.SizeMismatch(2, 3), op
    with pytest.raises(E) as e:
        V.from_slice([]).covariance(V.from_slice([]))                             # :1064
    assert e.value == E.EmptyVector
    with pytest.raises(E) as e:                                                   # src/matrix.rs:2010-2017
        M.from_vec(3, 2, [1, 2, 3, 4, 5, 6]).embedding_lookup([0, 5, 1])
    assert e.value == E.InvalidInput("Index 5 at position 1 is out of bounds for embedding table with 3 rows")


def test_widened_api_fails_loudly_without_gpu(trn):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-GPU contract is checked on CPU boxes")
    V, M = trn.Vector, trn.Matrix
    calls = [lambda: V.from_slice([1, 2, 3]).hardswish(), lambda: V.from_slice([1, 2, 3]).leaky_relu(0.1),
             lambda: V.from_slice([1, 2, 3]).zscore(), lambda: V.from_slice([1, 2, 3]).covariance(V.from_slice([3, 2, 1])),
             lambda: V.from_slice([1, 2, 3]).minimum(V.from_slice([3, 2, 1])), lambda: V.from_slice([1, 2, 3]).pow(2.0),
             lambda: M.from_vec(2, 2, [1, 2, 3, 4]).embedding_lookup([1, 0]),
             lambda: trn.softmax_rows(np.ones((2, 50257), np.float32), 2, 50257)]
    for c in calls:
        with pytest.raises(trn.TruenoError) as e:
            c()
        assert e.value.variant == "GpuError"
