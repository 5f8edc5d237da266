"""trueno_b200 — B200 (sm_100a) backend for paiml/trueno's data-parallel hot path.

This package is the Python view of the C-ABI library (include/trueno_cuda.h,
trueno_b200/libtrueno_cuda.so) plus a host-side mirror of the reference's public API for this
path — `Vector`, `Matrix`, `TruenoError` with the reference's names, argument meaning and
error behaviour (paiml/trueno src/vector.rs, src/matrix.rs, src/error.rs) — so that the parity
tests read like the reference's own tests.  The Rust binding a trueno maintainer would add is in
INTEGRATION.md; the C++ mirror is include/trueno.hpp.

There is NO CPU fallback anywhere in this package: if the CUDA library is missing the import
fails, and if no B200 is present every compute call raises TruenoError(GpuError).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libtrueno_cuda.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build it with `python trueno_b200/build.py` (nvcc, sm_100a). "
        "trueno_b200 has no CPU or PyTorch fallback."
    )

lib = C.CDLL(LIB_PATH)

_f32p = C.POINTER(C.c_float)
_u64p = C.POINTER(C.c_uint64)
_sz = C.c_size_t
_vp = C.c_void_p

_UNARY_EXT = ("abs", "relu", "exp", "swish", "tanh", "sqrt", "recip", "ln", "log2", "log10", "sin", "cos", "tan",
              "floor", "ceil", "round")
_REDUCE_EXT = ("sum_kahan", "norm_l1", "norm_linf", "mean", "variance", "stddev")
# the rest of Vector's element-wise API (src/vector.rs:2409-4410)
_UNARY_EXT2 = ("neg", "signum", "trunc", "fract", "sinh", "cosh", "asin", "acos", "atan", "asinh", "acosh", "atanh",
               "hardswish", "mish", "selu")
_BINARY_EXT2 = ("minimum", "maximum", "copysign")

# name -> argtypes; every function returns int (trn_status) unless listed in _RESTYPES
_SIGNATURES = {
    "trn_cuda_init": [C.c_int], "trn_cuda_shutdown": [], "trn_cuda_is_available": [],
    "trn_device_count": [C.POINTER(C.c_int)],
    "trn_device_info": [C.c_char_p, _sz, C.POINTER(C.c_int), _u64p],
    "trn_last_error": [C.c_char_p, _sz], "trn_last_mismatch": [_u64p, _u64p],
    "trn_synchronize": [_vp], "trn_launch_count": [],
    "trn_buf_alloc": [_sz, C.POINTER(_vp)], "trn_buf_free": [_vp],
    "trn_buf_upload": [_vp, _vp, _sz], "trn_buf_download": [_vp, _vp, _sz],
    "trn_buf_len": [_vp], "trn_buf_ptr": [_vp],
    "trn_host_alloc": [_sz, C.POINTER(_vp)], "trn_host_free": [_vp],
    "trn_dot_f32": [_vp, _sz, _vp, _sz, _f32p], "trn_sum_f32": [_vp, _sz, _f32p],
    "trn_max_f32": [_vp, _sz, _f32p], "trn_min_f32": [_vp, _sz, _f32p],
    "trn_argmax_f32": [_vp, _sz, _u64p], "trn_argmin_f32": [_vp, _sz, _u64p],
    "trn_norm_l2_f32": [_vp, _sz, _f32p],
    "trn_add_f32": [_vp, _sz, _vp, _sz, _vp], "trn_mul_f32": [_vp, _sz, _vp, _sz, _vp],
    "trn_sigmoid_f32": [_vp, _sz, _vp], "trn_gelu_f32": [_vp, _sz, _vp],
    "trn_softmax_rows_f32": [_vp, _vp, _sz, _sz], "trn_log_softmax_rows_f32": [_vp, _vp, _sz, _sz],
    "trn_matmul_f32": [_vp, _sz, _sz, _vp, _sz, _sz, _vp],
    "trn_batched_matmul_f32": [_vp, _sz, _vp, _sz, _vp, _sz, _sz, _sz, _sz],
    "trn_batched_matmul_4d_f32": [_vp, _sz, _vp, _sz, _vp, _sz, _sz, _sz, _sz, _sz],
    "trn_matvec_f32": [_vp, _sz, _sz, _vp, _sz, _vp],
    "trn_transpose_f32": [_vp, _sz, _sz, _vp],
    "trn_dot_f32_dev": [_vp, _sz, _vp, _sz, _vp, _vp], "trn_sum_f32_dev": [_vp, _sz, _vp, _vp],
    "trn_max_f32_dev": [_vp, _sz, _vp, _vp], "trn_min_f32_dev": [_vp, _sz, _vp, _vp],
    "trn_argmax_f32_dev": [_vp, _sz, _vp, _vp, _vp], "trn_argmin_f32_dev": [_vp, _sz, _vp, _vp, _vp],
    "trn_argmax_slice_f32_dev": [_vp, _sz, C.c_int, _vp, _vp, _vp],
    "trn_argmin_slice_f32_dev": [_vp, _sz, C.c_int, _vp, _vp, _vp],
    "trn_argmax_slice_pair_f32_dev": [_vp, _sz, C.c_uint64, _vp, _vp],
    "trn_argmin_slice_pair_f32_dev": [_vp, _sz, C.c_uint64, _vp, _vp],
    "trn_arg_combine_f32_dev": [_vp, _sz, C.c_int, _vp, _vp, _vp],
    "trn_norm_l2_f32_dev": [_vp, _sz, _vp, _vp], "trn_sumsq_f32_dev": [_vp, _sz, _vp, _vp],
    "trn_add_f32_dev": [_vp, _sz, _vp, _sz, _vp, _vp], "trn_mul_f32_dev": [_vp, _sz, _vp, _sz, _vp, _vp],
    "trn_sigmoid_f32_dev": [_vp, _sz, _vp, _vp], "trn_gelu_f32_dev": [_vp, _sz, _vp, _vp],
    "trn_softmax_rows_f32_dev": [_vp, _vp, _sz, _sz, _vp], "trn_log_softmax_rows_f32_dev": [_vp, _vp, _sz, _sz, _vp],
    "trn_matmul_f32_dev": [_vp, _sz, _sz, _vp, _sz, _sz, _vp, _vp],
    "trn_batched_matmul_f32_dev": [_vp, _sz, _vp, _sz, _vp, _sz, _sz, _sz, _sz, _vp],
    "trn_batched_matmul_4d_f32_dev": [_vp, _sz, _vp, _sz, _vp, _sz, _sz, _sz, _sz, _sz, _vp],
    "trn_matvec_f32_dev": [_vp, _sz, _sz, _vp, _sz, _vp, _vp],
    "trn_transpose_f32_dev": [_vp, _sz, _sz, _vp, _vp],
    **{f"trn_{n}_f32": [_vp, _sz, _vp] for n in _UNARY_EXT},
    **{f"trn_{n}_f32_dev": [_vp, _sz, _vp, _vp] for n in _UNARY_EXT},
    **{f"trn_{n}_f32": [_vp, _sz, _f32p] for n in _REDUCE_EXT},
    **{f"trn_{n}_f32_dev": [_vp, _sz, _vp, _vp] for n in _REDUCE_EXT[:3]},
    **{f"trn_{n}_f32": [_vp, _sz, _vp] for n in _UNARY_EXT2},
    **{f"trn_{n}_f32_dev": [_vp, _sz, _vp, _vp] for n in _UNARY_EXT2},
    **{f"trn_{n}_f32": [_vp, _sz, _vp, _sz, _vp] for n in _BINARY_EXT2},
    **{f"trn_{n}_f32_dev": [_vp, _sz, _vp, _sz, _vp, _vp] for n in _BINARY_EXT2},
    **{f"trn_{n}_f32": [_vp, _sz, C.c_float, _vp] for n in ("leaky_relu", "elu", "pow")},
    **{f"trn_{n}_f32_dev": [_vp, _sz, C.c_float, _vp, _vp] for n in ("leaky_relu", "elu", "pow")},
    "trn_clip_f32": [_vp, _sz, C.c_float, C.c_float, _vp], "trn_clip_f32_dev": [_vp, _sz, C.c_float, C.c_float, _vp, _vp],
    "trn_affine_f32_dev": [_vp, _sz, C.c_float, C.c_float, _vp, _vp],
    "trn_sum_of_squares_f32": [_vp, _sz, _f32p],
    "trn_covariance_f32": [_vp, _sz, _vp, _sz, _f32p], "trn_correlation_f32": [_vp, _sz, _vp, _sz, _f32p],
    "trn_zscore_f32": [_vp, _sz, _vp], "trn_minmax_normalize_f32": [_vp, _sz, _vp],
    "trn_layer_norm_simple_rows_f32": [_vp, C.c_float, _vp, _sz, _sz],
    "trn_layer_norm_simple_rows_f32_dev": [_vp, C.c_float, _vp, _sz, _sz, _vp],
    "trn_sub_f32": [_vp, _sz, _vp, _sz, _vp], "trn_sub_f32_dev": [_vp, _sz, _vp, _sz, _vp, _vp],
    "trn_div_f32": [_vp, _sz, _vp, _sz, _vp], "trn_div_f32_dev": [_vp, _sz, _vp, _sz, _vp, _vp],
    "trn_scale_f32": [_vp, _sz, C.c_float, _vp], "trn_scale_f32_dev": [_vp, _sz, C.c_float, _vp, _vp],
    "trn_clamp_f32": [_vp, _sz, C.c_float, C.c_float, _vp], "trn_clamp_f32_dev": [_vp, _sz, C.c_float, C.c_float, _vp, _vp],
    "trn_lerp_f32": [_vp, _sz, _vp, _sz, C.c_float, _vp], "trn_lerp_f32_dev": [_vp, _sz, _vp, _sz, C.c_float, _vp, _vp],
    "trn_fma_f32": [_vp, _sz, _vp, _sz, _vp, _sz, _vp], "trn_fma_f32_dev": [_vp, _sz, _vp, _sz, _vp, _sz, _vp, _vp],
    "trn_batch_create": [C.POINTER(_vp)], "trn_batch_destroy": [_vp],
    "trn_batch_upload": [_vp, _vp, _sz, C.POINTER(C.c_uint32)], "trn_batch_update": [_vp, C.c_uint32, _vp, _sz],
    "trn_batch_op": [_vp, C.c_int, C.c_uint32, C.c_uint32, C.c_float, C.POINTER(C.c_uint32)],
    "trn_batch_execute": [_vp], "trn_batch_read": [_vp, C.c_uint32, _vp, _sz],
    "trn_batch_num_operations": [_vp], "trn_batch_num_buffers": [_vp],
    "trn_vecmat_f32": [_vp, _sz, _vp, _sz, _sz, _vp], "trn_vecmat_f32_dev": [_vp, _sz, _vp, _sz, _sz, _vp, _vp],
    "trn_layer_norm_rows_f32": [_vp, _vp, _sz, _vp, _sz, C.c_float, _vp, _sz, _sz],
    "trn_layer_norm_rows_f32_dev": [_vp, _vp, _sz, _vp, _sz, C.c_float, _vp, _sz, _sz, _vp],
    "trn_comm_local_handle": [_vp], "trn_comm_create": [C.c_int, C.c_int, _vp, C.POINTER(_vp)], "trn_comm_destroy": [_vp],
    "trn_sum_allreduce_f32_dev": [_vp, _vp, _sz, _vp, _vp], "trn_dot_allreduce_f32_dev": [_vp, _vp, _sz, _vp, _sz, _vp, _vp],
    "trn_norm_l2_allreduce_f32_dev": [_vp, _vp, _sz, _vp, _vp],
    "trn_argmax_allgather_f32_dev": [_vp, _vp, _sz, C.c_uint64, _vp, _vp, _vp],
    "trn_argmin_allgather_f32_dev": [_vp, _vp, _sz, C.c_uint64, _vp, _vp, _vp],
    "trn_softmax_slice_stats_f32_dev": [_vp, _sz, _vp, _vp],
    "trn_softmax_slice_apply_f32_dev": [_vp, _sz, _vp, _sz, C.c_int, _vp, _vp],
    "trn_embedding_lookup_f32": [_vp, _sz, _sz, _vp, _sz, _vp], "trn_embedding_lookup_f32_dev": [_vp, _sz, _sz, _vp, _sz, _vp, _vp],
    "trn_convolve2d_f32": [_vp, _sz, _sz, _vp, _sz, _sz, _vp], "trn_convolve2d_f32_dev": [_vp, _sz, _sz, _vp, _sz, _sz, _vp, _vp],
    "trn_attention_f32": [_vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _sz, _sz, C.c_float, C.c_int],
    "trn_attention_f32_dev": [_vp, _sz, _vp, _sz, _vp, _sz, _vp, _sz, _sz, _sz, C.c_float, C.c_int, _vp],
    "trn_symmetric_eigen_f32": [_vp, _sz, _sz, _vp, _vp], "trn_symmetric_eigen_f32_dev": [_vp, _sz, _sz, _vp, _vp, _vp],
    "trn_comm_status": [_vp],
    "trn_gemm_prepare_b_dev": [_vp, _sz, _sz, C.POINTER(_vp), _vp], "trn_gemm_b_free": [_vp],
    "trn_matmul_prepared_f32_dev": [_vp, _sz, _sz, _vp, _vp, _vp],
    "trn_matmul_prepared_f32": [_vp, _sz, _sz, _vp, _vp],
    "trn_matmul_rowblock_f32_dev": [_vp, _sz, _sz, _sz, _vp, _sz, _sz, _vp, _vp],
    "trn_matmul_rowblock_prepared_f32_dev": [_vp, _sz, _sz, _sz, _vp, _vp, _vp],
    "trn_set_gemm_engine": [C.c_int], "trn_get_gemm_engine": [],
    "trn_profile_enable": [C.c_int], "trn_profile_last_gemm": [_f32p, _f32p],
}
_RESTYPES = {"trn_last_error": _sz, "trn_last_mismatch": None, "trn_launch_count": C.c_uint64,
             "trn_buf_len": _sz, "trn_buf_ptr": _vp, "trn_batch_num_operations": _sz, "trn_batch_num_buffers": _sz}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
for _name, _args in _SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError here == the library does not export what the header declares
    _fn.argtypes = _args
    _fn.restype = _RESTYPES.get(_name, C.c_int)

ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC_3XTF32, ENGINE_TC_1XTF32 = 0, 1, 2, 3

_VARIANTS = {1: "SizeMismatch", 2: "InvalidInput", 3: "EmptyVector", 4: "DivisionByZero", 5: "GpuError",
             6: "UnsupportedBackend"}


class TruenoError(Exception):
    """Mirror of `enum TruenoError` (src/error.rs:8-41).  Compares equal by variant + payload, as the
    reference's tests compare error values (src/vector.rs:4853-4856)."""

    def __init__(self, variant: str, message: str = "", expected: int | None = None, actual: int | None = None):
        self.variant, self.message, self.expected, self.actual = variant, message, expected, actual
        super().__init__(self.__str__())

    def __str__(self):
        if self.variant == "SizeMismatch":
            return f"Size mismatch: expected {self.expected}, got {self.actual}"
        if self.variant == "InvalidInput":
            return f"Invalid input: {self.message}"
        if self.variant == "GpuError":
            return f"GPU error: {self.message}"
        if self.variant == "EmptyVector":
            return "Empty vector"
        if self.variant == "DivisionByZero":
            return "Division by zero"
        return f"Backend not supported on this platform: {self.message}"

    def __eq__(self, other):
        return (isinstance(other, TruenoError) and self.variant == other.variant and
                (self.variant in ("EmptyVector", "DivisionByZero") or
                 (self.variant == "SizeMismatch" and (self.expected, self.actual) == (other.expected, other.actual)) or
                 (self.variant not in ("SizeMismatch",) and self.message == other.message)))

    __hash__ = Exception.__hash__

    @staticmethod
    def InvalidInput(msg: str) -> "TruenoError":
        return TruenoError("InvalidInput", msg)

    @staticmethod
    def SizeMismatch(expected: int, actual: int) -> "TruenoError":
        return TruenoError("SizeMismatch", "", expected, actual)

    EmptyVector: "TruenoError"


TruenoError.EmptyVector = TruenoError("EmptyVector")


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    lib.trn_last_error(buf, 1024)
    return buf.value.decode("utf-8", "replace")


def check(status: int) -> None:
    """Turns a trn_status into the TruenoError the reference would return."""
    if status == 0:
        return
    variant = _VARIANTS.get(status, "GpuError")
    if variant == "SizeMismatch":
        e, a = C.c_uint64(), C.c_uint64()
        lib.trn_last_mismatch(C.byref(e), C.byref(a))
        raise TruenoError(variant, "", e.value, a.value)
    raise TruenoError(variant, last_error())


def is_available() -> bool:
    return bool(lib.trn_cuda_is_available())


def device_info() -> dict:
    name = C.create_string_buffer(256)
    sms, hbm = C.c_int(), C.c_uint64()
    check(lib.trn_device_info(name, 256, C.byref(sms), C.byref(hbm)))
    return {"name": name.value.decode(), "sm_count": sms.value, "hbm_bytes": hbm.value}


def set_gemm_engine(engine: int) -> None:
    check(lib.trn_set_gemm_engine(engine))


def launch_count() -> int:
    return int(lib.trn_launch_count())


def synchronize(stream: int | None = None) -> None:
    check(lib.trn_synchronize(stream))


def _as_f32(data) -> np.ndarray:
    return np.ascontiguousarray(data, dtype=np.float32).reshape(-1)


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def pinned_empty(n: int) -> np.ndarray:
    """numpy view of page-locked host memory from trn_host_alloc (freed when the array is collected)."""
    p = _vp()
    check(lib.trn_host_alloc(n, C.byref(p)))
    buf = (C.c_float * max(n, 1)).from_address(p.value)
    arr = np.frombuffer(buf, dtype=np.float32, count=n)

    class _Owner:
        def __init__(self, addr): self.addr = addr
        def __del__(self):
            try:
                lib.trn_host_free(self.addr)
            except Exception:
                pass
    arr = arr.view(_PinnedArray)
    arr._owner = _Owner(p.value)
    return arr


class _PinnedArray(np.ndarray):
    _owner = None

    def __array_finalize__(self, obj):
        if obj is not None:
            self._owner = getattr(obj, "_owner", None)


class DeviceBuffer:
    """The device-buffer type of the north star: f32 storage resident in HBM with pinned staging
    (trn_buf_*; precedent: GpuCommandBatch::upload/read, src/backends/gpu/batch.rs:140-200)."""

    def __init__(self, n: int):
        h = _vp()
        check(lib.trn_buf_alloc(n, C.byref(h)))
        self._h, self.len = h, n

    @classmethod
    def from_host(cls, data) -> "DeviceBuffer":
        a = _as_f32(data)
        b = cls(a.size)
        b.upload(a)
        return b

    @property
    def ptr(self) -> int:
        return lib.trn_buf_ptr(self._h) or 0

    def upload(self, data) -> None:
        a = _as_f32(data)
        check(lib.trn_buf_upload(self._h, _ptr(a), a.size))

    def to_host(self) -> np.ndarray:
        out = np.empty(self.len, np.float32)
        check(lib.trn_buf_download(self._h, _ptr(out), out.size))
        return out

    def free(self) -> None:
        if self._h is not None:
            lib.trn_buf_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Vector:
    """Mirror of `Vector<f32>` (src/vector.rs:125-128) for the hot-path ops.  Host data in, host
    data out; every op runs on the B200 through the C ABI."""

    def __init__(self, data: np.ndarray):
        self.data = data

    @staticmethod
    def from_slice(data) -> "Vector":
        return Vector(_as_f32(data).copy())

    from_vec = from_slice

    def as_slice(self) -> np.ndarray:
        return self.data

    def len(self) -> int:
        return int(self.data.size)

    def __len__(self):
        return self.len()

    # ---- reductions (src/vector.rs:588-827, 2601) ----
    def dot(self, other: "Vector") -> np.float32:
        out = C.c_float()
        check(lib.trn_dot_f32(_ptr(self.data), self.data.size, _ptr(other.data), other.data.size, C.byref(out)))
        return np.float32(out.value)

    def _reduce(self, fn) -> np.float32:
        out = C.c_float()
        check(fn(_ptr(self.data), self.data.size, C.byref(out)))
        return np.float32(out.value)

    def sum(self): return self._reduce(lib.trn_sum_f32)
    def max(self): return self._reduce(lib.trn_max_f32)
    def min(self): return self._reduce(lib.trn_min_f32)
    def norm_l2(self): return self._reduce(lib.trn_norm_l2_f32)

    def _arg(self, fn) -> int:
        out = C.c_uint64()
        check(fn(_ptr(self.data), self.data.size, C.byref(out)))
        return int(out.value)

    def argmax(self): return self._arg(lib.trn_argmax_f32)
    def argmin(self): return self._arg(lib.trn_argmin_f32)

    # ---- elementwise (src/vector.rs:358, 478, 1854, 2179) ----
    def _binary(self, fn, other: "Vector") -> "Vector":
        out = np.empty(self.data.size, np.float32)
        check(fn(_ptr(self.data), self.data.size, _ptr(other.data), other.data.size, _ptr(out)))
        return Vector(out)

    def add(self, other): return self._binary(lib.trn_add_f32, other)
    def mul(self, other): return self._binary(lib.trn_mul_f32, other)

    def _map(self, fn) -> "Vector":
        out = np.empty(self.data.size, np.float32)
        check(fn(_ptr(self.data), self.data.size, _ptr(out)))
        return Vector(out)

    def sigmoid(self): return self._map(lib.trn_sigmoid_f32)
    def gelu(self): return self._map(lib.trn_gelu_f32)

    # ---- remaining VectorBackend surface (src/vector.rs:423-4182) ----
    def sub(self, other): return self._binary(lib.trn_sub_f32, other)
    def div(self, other): return self._binary(lib.trn_div_f32, other)

    def scale(self, scalar: float) -> "Vector":
        out = np.empty(self.data.size, np.float32)
        check(lib.trn_scale_f32(_ptr(self.data), self.data.size, scalar, _ptr(out)))
        return Vector(out)

    def clamp(self, min_val: float, max_val: float) -> "Vector":
        out = np.empty(self.data.size, np.float32)
        check(lib.trn_clamp_f32(_ptr(self.data), self.data.size, min_val, max_val, _ptr(out)))
        return Vector(out)

    def lerp(self, other: "Vector", t: float) -> "Vector":
        out = np.empty(self.data.size, np.float32)
        check(lib.trn_lerp_f32(_ptr(self.data), self.data.size, _ptr(other.data), other.data.size, t, _ptr(out)))
        return Vector(out)

    def fma(self, b: "Vector", c: "Vector") -> "Vector":
        out = np.empty(self.data.size, np.float32)
        check(lib.trn_fma_f32(_ptr(self.data), self.data.size, _ptr(b.data), b.data.size, _ptr(c.data), c.data.size, _ptr(out)))
        return Vector(out)

    # ---- the rest of Vector's element-wise / statistics API (unary maps are installed below) ----
    def _map1(self, fn, *params) -> "Vector":
        out = np.empty(self.data.size, np.float32)
        check(fn(_ptr(self.data), self.data.size, *params, _ptr(out)))
        return Vector(out)

    def leaky_relu(self, negative_slope: float): return self._map1(lib.trn_leaky_relu_f32, negative_slope)   # src/vector.rs:1980
    def elu(self, alpha: float): return self._map1(lib.trn_elu_f32, alpha)                                   # :2085
    def pow(self, n: float): return self._map1(lib.trn_pow_f32, n)                                           # :3342
    def clip(self, min_val: float, max_val: float): return self._map1(lib.trn_clip_f32, min_val, max_val)    # :1448
    def minimum(self, other): return self._binary(lib.trn_minimum_f32, other)                                # :4328
    def maximum(self, other): return self._binary(lib.trn_maximum_f32, other)                                # :4364
    def copysign(self, sign): return self._binary(lib.trn_copysign_f32, sign)                                # :4292
    def zscore(self): return self._map(lib.trn_zscore_f32)                                                   # :1180
    def minmax_normalize(self): return self._map(lib.trn_minmax_normalize_f32)                               # :1248
    def sum_of_squares(self): return self._reduce(lib.trn_sum_of_squares_f32)                                # :898

    def _reduce2(self, fn, other: "Vector") -> float:
        out = C.c_float()
        check(fn(_ptr(self.data), self.data.size, _ptr(other.data), other.data.size, C.byref(out)))
        return float(np.float32(out.value))

    def covariance(self, other): return self._reduce2(lib.trn_covariance_f32, other)                         # :1063
    def correlation(self, other): return self._reduce2(lib.trn_correlation_f32, other)                       # :1119

    def layer_norm_simple(self, eps: float) -> "Vector":                                                     # :1386
        out = np.empty(self.data.size, np.float32)
        check(lib.trn_layer_norm_simple_rows_f32(_ptr(self.data), eps, _ptr(out), 1, self.data.size))
        return Vector(out)

    def sum_kahan(self): return self._reduce(lib.trn_sum_kahan_f32)
    def norm_l1(self): return self._reduce(lib.trn_norm_l1_f32)
    def norm_linf(self): return self._reduce(lib.trn_norm_linf_f32)
    def mean(self): return self._reduce(lib.trn_mean_f32)
    def variance(self): return self._reduce(lib.trn_variance_f32)
    def stddev(self): return self._reduce(lib.trn_stddev_f32)

    def normalize(self) -> "Vector":
        """src/vector.rs:2665-2678: norm_l2, |norm| < 1e-10 -> DivisionByZero, else scale(1 / norm)."""
        norm = self.norm_l2()
        if abs(float(norm)) < 1e-10:
            raise TruenoError("DivisionByZero")
        return self.scale(float(np.float32(1.0) / norm))

    def layer_norm(self, gamma: "Vector", beta: "Vector", eps: float) -> "Vector":
        """Vector::layer_norm (src/vector.rs:1316)."""
        out = np.empty(self.data.size, np.float32)
        check(lib.trn_layer_norm_rows_f32(_ptr(self.data), _ptr(gamma.data), gamma.data.size, _ptr(beta.data), beta.data.size,
                                          eps, _ptr(out), 1 if self.data.size else 0, self.data.size))
        return Vector(out)

    # ---- softmax family (src/vector.rs:1516, 1581): a Vector is one row ----
    def softmax(self) -> "Vector":
        out = np.empty(self.data.size, np.float32)
        check(lib.trn_softmax_rows_f32(_ptr(self.data), _ptr(out), 1 if self.data.size else 0, self.data.size))
        return Vector(out)

    def log_softmax(self) -> "Vector":
        out = np.empty(self.data.size, np.float32)
        check(lib.trn_log_softmax_rows_f32(_ptr(self.data), _ptr(out), 1 if self.data.size else 0, self.data.size))
        return Vector(out)


class Matrix:
    """Mirror of `Matrix<f32>` (src/matrix.rs:49-54): dense row-major."""

    def __init__(self, rows: int, cols: int, data: np.ndarray):
        self._rows, self._cols, self.data = rows, cols, data

    @staticmethod
    def from_vec(rows: int, cols: int, data) -> "Matrix":
        a = _as_f32(data)
        if a.size != rows * cols:  # src/matrix.rs:108-117
            raise TruenoError.InvalidInput(
                f"Data length {a.size} does not match matrix dimensions {rows}x{cols} (expected {rows * cols})")
        return Matrix(rows, cols, a.copy())

    from_slice = from_vec

    @staticmethod
    def zeros(rows: int, cols: int) -> "Matrix":
        return Matrix(rows, cols, np.zeros(rows * cols, np.float32))

    @staticmethod
    def identity(n: int) -> "Matrix":
        return Matrix(n, n, np.eye(n, dtype=np.float32).reshape(-1))

    def rows(self): return self._rows
    def cols(self): return self._cols
    def shape(self): return (self._rows, self._cols)
    def as_slice(self): return self.data

    def get(self, i: int, j: int):
        return self.data[i * self._cols + j] if i < self._rows and j < self._cols else None

    def to_numpy(self) -> np.ndarray:
        return self.data.reshape(self._rows, self._cols)

    def matmul(self, other: "Matrix") -> "Matrix":
        out = np.empty(self._rows * other._cols if self._cols == other._rows else 0, np.float32)
        check(lib.trn_matmul_f32(_ptr(self.data), self._rows, self._cols, _ptr(other.data), other._rows, other._cols,
                                 _ptr(out)))
        return Matrix(self._rows, other._cols, out)

    def matvec(self, v: Vector) -> Vector:
        out = np.empty(self._rows, np.float32)
        check(lib.trn_matvec_f32(_ptr(self.data), self._rows, self._cols, _ptr(v.data), v.data.size, _ptr(out)))
        return Vector(out)

    def transpose(self) -> "Matrix":
        out = np.empty(self.data.size, np.float32)
        check(lib.trn_transpose_f32(_ptr(self.data), self._rows, self._cols, _ptr(out)))
        return Matrix(self._cols, self._rows, out)

    def convolve2d(self, kernel: "Matrix") -> "Matrix":
        """Matrix::convolve2d (src/matrix.rs:1868): valid-padding cross-correlation."""
        ok = kernel._rows <= self._rows and kernel._cols <= self._cols
        orows, ocols = (self._rows - kernel._rows + 1, self._cols - kernel._cols + 1) if ok else (0, 0)
        out = np.empty(orows * ocols, np.float32)
        check(lib.trn_convolve2d_f32(_ptr(self.data), self._rows, self._cols, _ptr(kernel.data), kernel._rows, kernel._cols, _ptr(out)))
        return Matrix(orows, ocols, out)

    def embedding_lookup(self, indices) -> "Matrix":
        """Matrix::embedding_lookup (src/matrix.rs:2008): rows of the table selected by `indices` (usize)."""
        idx = np.ascontiguousarray(np.asarray(indices, dtype=np.uint64).reshape(-1))
        out = np.empty(idx.size * self._cols, np.float32)
        check(lib.trn_embedding_lookup_f32(_ptr(self.data), self._rows, self._cols, idx.ctypes.data, idx.size, _ptr(out)))
        return Matrix(idx.size, self._cols, out)

    def embedding_lookup_sparse(self, indices):
        """Matrix::embedding_lookup_sparse (src/matrix.rs:2059): (embeddings, sorted unique indices)."""
        emb = self.embedding_lookup(indices)
        return emb, sorted(set(int(i) for i in np.asarray(indices, dtype=np.uint64).reshape(-1)))

    @staticmethod
    def vecmat(v: Vector, m: "Matrix") -> Vector:
        """Matrix::vecmat (src/matrix.rs:1782): v^T * m."""
        out = np.empty(m._cols, np.float32)
        check(lib.trn_vecmat_f32(_ptr(v.data), v.data.size, _ptr(m.data), m._rows, m._cols, _ptr(out)))
        return Vector(out)

    @staticmethod
    def batched_matmul(a, b, batch: int, m: int, k: int, n: int) -> np.ndarray:
        a, b = _as_f32(a), _as_f32(b)
        out = np.empty(batch * m * n, np.float32)
        check(lib.trn_batched_matmul_f32(_ptr(a), a.size, _ptr(b), b.size, _ptr(out), batch, m, k, n))
        return out

    @staticmethod
    def batched_matmul_4d(a, b, batch: int, heads: int, m: int, k: int, n: int) -> np.ndarray:
        a, b = _as_f32(a), _as_f32(b)
        out = np.empty(batch * heads * m * n, np.float32)
        check(lib.trn_batched_matmul_4d_f32(_ptr(a), a.size, _ptr(b), b.size, _ptr(out), batch, heads, m, k, n))
        return out


class SymmetricEigen:
    """Mirror of `SymmetricEigen` (src/eigen.rs:56-516): eigenvalues in descending order, eigenvectors as the columns
    of a Matrix, computed by parallel Jacobi rotations on the device (`trn_symmetric_eigen_f32`)."""

    def __init__(self, matrix: "Matrix"):
        n = matrix.rows()
        vals = np.empty(n if matrix.rows() == matrix.cols() else 0, np.float32)
        vecs = np.empty(vals.size * vals.size, np.float32)
        check(lib.trn_symmetric_eigen_f32(_ptr(matrix.data), matrix.rows(), matrix.cols(), _ptr(vals), _ptr(vecs)))
        self._values = vals
        self._vectors = Matrix(n, n, vecs)

    @staticmethod
    def new(matrix: "Matrix") -> "SymmetricEigen":
        return SymmetricEigen(matrix)

    def eigenvalues(self) -> np.ndarray:
        return self._values

    def eigenvectors(self) -> "Matrix":
        return self._vectors

    def __len__(self) -> int:
        return int(self._values.size)

    def is_empty(self) -> bool:
        return self._values.size == 0

    def eigenvector(self, i: int):
        """src/eigen.rs:434: column i as a Vector, None when out of range."""
        n = len(self)
        if not 0 <= i < n:
            return None
        return Vector.from_slice(self._vectors.data.reshape(n, n)[:, i].copy())

    def __iter__(self):
        """src/eigen.rs:403: (eigenvalue, eigenvector) pairs in descending order."""
        return ((float(self._values[i]), self.eigenvector(i)) for i in range(len(self)))

    def reconstruct(self) -> "Matrix":
        """src/eigen.rs:468: V * diag(lambda) * V^T on the device GEMM."""
        n = len(self)
        v = self._vectors.data.reshape(n, n)
        scaled = Matrix(n, n, (v * self._values[None, :]).astype(np.float32).ravel())
        return scaled.matmul(self._vectors.transpose())


def attention(q, k, v, heads: int, seq_len: int, head_dim: int, scale: float | None = None, causal: bool = False) -> np.ndarray:
    """Fused attention out = softmax(scale * Q K^T [causal]) V per head; q, k, v are [heads][seq_len][head_dim].
    Mirrors trueno-gpu's `AttentionKernel::new(seq_len, head_dim)[.with_causal()][.with_scale(s)]`
    (trueno-gpu/src/kernels/attention.rs:46-112): `scale` defaults to 1/sqrt(head_dim)."""
    q, k, v = _as_f32(q), _as_f32(k), _as_f32(v)
    if scale is None:
        scale = 1.0 / float(np.sqrt(np.float32(head_dim))) if head_dim else 1.0
    out = np.empty(heads * seq_len * head_dim, np.float32)
    check(lib.trn_attention_f32(_ptr(q), q.size, _ptr(k), k.size, _ptr(v), v.size, _ptr(out), heads, seq_len, head_dim,
                                float(scale), int(bool(causal))))
    return out


class CommandBatch:
    """Mirror of `GpuCommandBatch` (src/backends/gpu/batch.rs:118-1019): queue uploads and ops, `execute()` them as
    one CUDA graph launch on the B200, `read()` results back.  BufferIds are plain ints."""
    _OPS = {"relu": 0, "scale": 1, "add": 2, "mul": 3, "dot": 4, "sigmoid": 5, "tanh": 6, "swish": 7, "gelu": 8, "sub": 9}

    def __init__(self):
        h = _vp()
        check(lib.trn_batch_create(C.byref(h)))
        self._h = h
        self._sizes: list[int] = []

    def upload(self, data) -> int:
        a = _as_f32(data)
        bid = C.c_uint32()
        check(lib.trn_batch_upload(self._h, _ptr(a), a.size, C.byref(bid)))
        self._sizes.append(a.size)
        return bid.value

    def update(self, buffer_id: int, data) -> None:
        a = _as_f32(data)
        check(lib.trn_batch_update(self._h, buffer_id, _ptr(a), a.size))

    def _op(self, name: str, a: int, b: int = 0, scalar: float = 0.0) -> int:
        out = C.c_uint32()
        check(lib.trn_batch_op(self._h, self._OPS[name], a, b, scalar, C.byref(out)))
        self._sizes.append(1 if name == "dot" else self._sizes[a])
        return out.value

    def relu(self, x): return self._op("relu", x)
    def scale(self, x, scalar: float): return self._op("scale", x, 0, scalar)
    def add(self, a, b): return self._op("add", a, b)
    def mul(self, a, b): return self._op("mul", a, b)
    def dot(self, a, b): return self._op("dot", a, b)
    def sigmoid(self, x): return self._op("sigmoid", x)
    def tanh(self, x): return self._op("tanh", x)
    def swish(self, x): return self._op("swish", x)
    def gelu(self, x): return self._op("gelu", x)
    def sub(self, a, b): return self._op("sub", a, b)

    def execute(self) -> None:
        check(lib.trn_batch_execute(self._h))

    def read(self, buffer_id: int) -> np.ndarray:
        n = self._sizes[buffer_id] if buffer_id < len(self._sizes) else 0
        out = np.empty(n, np.float32)
        check(lib.trn_batch_read(self._h, buffer_id, _ptr(out), n))
        return out

    def num_operations(self) -> int: return int(lib.trn_batch_num_operations(self._h))
    def num_buffers(self) -> int: return int(lib.trn_batch_num_buffers(self._h))

    def __del__(self):
        try:
            if self._h is not None:
                lib.trn_batch_destroy(self._h)
                self._h = None
        except Exception:
            pass


def softmax_rows(a, rows: int, cols: int, log: bool = False) -> np.ndarray:
    """`rows` independent Vector::softmax calls in one launch (GpuDevice::softmax hook, device.rs:951)."""
    a = _as_f32(a)
    out = np.empty(a.size, np.float32)
    fn = lib.trn_log_softmax_rows_f32 if log else lib.trn_softmax_rows_f32
    check(fn(_ptr(a), _ptr(out), rows, cols))
    return out.reshape(rows, cols)


def _install_unary_maps():
    def make(name):
        fn = getattr(lib, f"trn_{name}_f32")

        def method(self):
            return self._map(fn)
        method.__name__ = name
        method.__doc__ = f"Vector::{name} (src/vector.rs) -> trn_{name}_f32"
        return method
    for name in _UNARY_EXT + _UNARY_EXT2:
        setattr(Vector, name, make(name))


_install_unary_maps()
