"""BASELINE.json's configurations at FULL size on one B200, checked through size-independent properties
and device-side f64 references on sampled rows (the CPU oracle cannot finish these sizes in seconds; its
parity at small sizes is tests/test_parity_gpu.py).  torch is used here as device memory and as an f64
checker only — every hot-path result comes from the C ABI `_dev` entry points.

Inputs: x = u01(splitmix64(seed ^ idx)), regenerated on the device."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def u01(seed: int, n: int, dev):
    idx = torch.arange(n, dtype=torch.int64, device=dev)
    z = (idx ^ seed) + (-7046029254386353131)
    z = (z ^ (z >> 30 & 0x3FFFFFFFF)) * (-4658895280553007687)
    z = (z ^ (z >> 27 & 0x1FFFFFFFFF)) * (-7723592293110705685)
    z = z ^ (z >> 31 & 0x1FFFFFFFF)
    return ((z >> 40) & 0xFFFFFF).to(torch.float32) * (1.0 / (1 << 24))


@pytest.fixture(scope="module")
def gpu(trn):
    trn.check(trn.lib.trn_cuda_init(0))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    return dev


def _st():
    return torch.cuda.current_stream().cuda_stream or 1


def _chunked_f64(fn, x, chunk=1 << 26):
    """sum over chunks of fn(chunk.double()) — keeps the f64 temporary at 512 MiB"""
    tot = torch.zeros((), dtype=torch.float64, device=x.device)
    for i in range(0, x.numel(), chunk):
        tot += fn(x[i:i + chunk].double())
    return float(tot)


# ---- config 2: Matrix::matmul 8192^3 -------------------------------------------------------------------
@pytest.mark.parametrize("signed", [False, True], ids=["u01", "u-11"])
def test_config2_matmul_8192(trn, gpu, signed):
    n = 8192
    a = u01(0x5EED0001, n * n, gpu).view(n, n)
    b = u01(0x5EED0002, n * n, gpu).view(n, n)
    if signed:
        a, b = a * 2 - 1, b * 2 - 1
    c = torch.empty(n, n, device=gpu)
    trn.check(trn.lib.trn_matmul_f32_dev(a.data_ptr(), n, n, b.data_ptr(), n, n, c.data_ptr(), _st()))
    torch.cuda.synchronize()
    # sampled rows against the f64 product, on the sum|a||b| scale (1e-5 contract)
    rows = torch.tensor([0, 1, 127, 128, 4095, 4096, 8191] + list(range(1000, 8000, 997)), device=gpu)
    truth = a[rows].double() @ b.double()
    scale = a[rows].double().abs() @ b.double().abs()
    err = ((c[rows].double() - truth).abs() / scale).max().item()
    assert err <= 1e-5, err
    # linearity / checksum of checksums: (A B) 1 == A (B 1) — touches EVERY output element
    ones = torch.ones(n, dtype=torch.float64, device=gpu)
    lhs = c.double() @ ones
    rhs = a.double() @ (b.double() @ ones)
    bound = a.double().abs() @ (b.double().abs() @ ones)
    assert ((lhs - rhs).abs() / bound).max().item() <= 1e-5
    # bit-identical rerun (tests/wasm_optimization_tests.rs:200-230)
    c2 = torch.empty_like(c)
    trn.check(trn.lib.trn_matmul_f32_dev(a.data_ptr(), n, n, b.data_ptr(), n, n, c2.data_ptr(), _st()))
    torch.cuda.synchronize()
    assert torch.equal(c, c2)


def test_config2_simt_vs_tensor_core_engines(trn, gpu):
    """BASELINE config 2: '3xTF32 tcgen05 vs SIMT FFMA, tolerance vs CPU' — both engines inside the same bound."""
    m, k, n = 1024, 8192, 1024
    a = u01(0x5EED0001, m * k, gpu).view(m, k) * 2 - 1
    b = u01(0x5EED0002, k * n, gpu).view(k, n) * 2 - 1
    truth = a.double() @ b.double()
    scale = a.double().abs() @ b.double().abs()
    errs = {}
    for name, eng in (("simt", trn.ENGINE_SIMT), ("tc3", trn.ENGINE_TC_3XTF32)):
        trn.set_gemm_engine(eng)
        c = torch.empty(m, n, device=gpu)
        trn.check(trn.lib.trn_matmul_f32_dev(a.data_ptr(), m, k, b.data_ptr(), k, n, c.data_ptr(), _st()))
        torch.cuda.synchronize()
        errs[name] = ((c.double() - truth).abs() / scale).max().item()
    trn.set_gemm_engine(trn.ENGINE_AUTO)
    assert errs["simt"] <= 1e-5 and errs["tc3"] <= 1e-5, errs


# ---- config 3: batched_matmul_4d B=8 H=32 m=2048 k=128 n=2048 ---------------------------------------------
def test_config3_batched_matmul_4d_full(trn, gpu):
    B, H, m, k, n = 8, 32, 2048, 128, 2048
    a = u01(0x5EED0003, B * H * m * k, gpu) * 2 - 1
    b = u01(0x5EED0004, B * H * k * n, gpu) * 2 - 1
    c = torch.empty(B * H * m * n, device=gpu)
    trn.check(trn.lib.trn_batched_matmul_4d_f32_dev(a.data_ptr(), a.numel(), b.data_ptr(), b.numel(), c.data_ptr(),
                                                    B, H, m, k, n, _st()))
    torch.cuda.synchronize()
    a3, b3, c3 = a.view(B * H, m, k), b.view(B * H, k, n), c.view(B * H, m, n)
    for h in (0, 1, 31, 32, 100, 255):
        truth = a3[h].double() @ b3[h].double()
        scale = a3[h].double().abs() @ b3[h].double().abs()
        assert ((c3[h].double() - truth).abs() / scale).max().item() <= 1e-5
        # head independence: the same head computed alone is bit-identical (src/matrix.rs:507-524 loops heads)
        solo = torch.empty(m, n, device=gpu)
        trn.check(trn.lib.trn_matmul_f32_dev(a3[h].data_ptr(), m, k, b3[h].data_ptr(), k, n, solo.data_ptr(), _st()))
        torch.cuda.synchronize()
        assert torch.equal(solo, c3[h])
    # every head's checksum: sum over the head of C == sum_k (colsum_k(A) * rowsum_k(B))
    lhs = c3.double().sum(dim=(1, 2))
    rhs = (a3.double().sum(dim=1) * b3.double().sum(dim=2)).sum(dim=1)
    bound = (a3.double().abs().sum(dim=1) * b3.double().abs().sum(dim=2)).sum(dim=1)
    assert ((lhs - rhs).abs() / bound).max().item() <= 1e-5


# ---- config 4: dot / sum / argmax / norm_l2 on 2^30 f32 ----------------------------------------------------
def test_config4_reductions_2pow30(trn, gpu):
    n = 1 << 30
    L = trn.lib
    x = u01(0x5EED0005, n, gpu) * 2 - 1
    y = u01(0x5EED0006, n, gpu) * 2 - 1
    out = torch.zeros(4, device=gpu)
    idx = torch.zeros(2, dtype=torch.int64, device=gpu)

    def scalar(fn, *args):
        trn.check(fn(*args, out.data_ptr(), _st()))
        torch.cuda.synchronize()
        return float(out[0])

    s = scalar(L.trn_sum_f32_dev, x.data_ptr(), n)
    assert abs(s - _chunked_f64(torch.sum, x)) <= 1e-5 * _chunked_f64(lambda t: t.abs().sum(), x)
    nrm = scalar(L.trn_norm_l2_f32_dev, x.data_ptr(), n)
    tn = _chunked_f64(lambda t: (t * t).sum(), x) ** 0.5
    assert abs(nrm - tn) <= 1e-5 * tn
    chunk = 1 << 26
    tdot = sum(float((x[i:i + chunk].double() * y[i:i + chunk].double()).sum()) for i in range(0, n, chunk))
    adot = sum(float((x[i:i + chunk].double() * y[i:i + chunk].double()).abs().sum()) for i in range(0, n, chunk))
    trn.check(L.trn_dot_f32_dev(x.data_ptr(), n, y.data_ptr(), n, out.data_ptr(), _st()))
    torch.cuda.synchronize()
    assert abs(float(out[0]) - tdot) <= 1e-5 * adot
    # checksum of checksums: the 8 contiguous slices of SURVEY.md §8e add up to the whole
    parts = []
    for r in range(8):
        sl = x[r * (n // 8):(r + 1) * (n // 8)]
        parts.append(scalar(L.trn_sum_f32_dev, sl.data_ptr(), sl.numel()))
    assert abs(sum(parts) - s) <= 1e-5 * _chunked_f64(lambda t: t.abs().sum(), x)

    # argmax / argmin need > 31-bit-safe indices and first-occurrence ties: plant equal maxima at two
    # far-apart indices (the second beyond 2^24 and 2^29), the lower one must win; same for minima
    hi_idx, hi_dup = (1 << 24) + 12345, (1 << 29) + 777
    lo_idx, lo_dup = (1 << 27) + 3, (1 << 30) - 5
    x[hi_idx] = 2.0; x[hi_dup] = 2.0
    x[lo_idx] = -3.0; x[lo_dup] = -3.0
    trn.check(L.trn_argmax_f32_dev(x.data_ptr(), n, idx.data_ptr(), out.data_ptr(), _st()))
    torch.cuda.synchronize()
    assert int(idx[0]) == hi_idx and float(out[0]) == 2.0
    trn.check(L.trn_argmin_f32_dev(x.data_ptr(), n, idx.data_ptr(), out.data_ptr(), _st()))
    torch.cuda.synchronize()
    assert int(idx[0]) == lo_idx and float(out[0]) == -3.0
    # idempotence of max: max(x) is an element of x and nothing exceeds it
    trn.check(L.trn_max_f32_dev(x.data_ptr(), n, out.data_ptr(), _st()))
    torch.cuda.synchronize()
    assert float(out[0]) == float(x.max())
    # NaN elements never win unless they sit at index 0 (src/backends/scalar.rs:140-166)
    x[5] = float("nan")
    trn.check(L.trn_argmax_f32_dev(x.data_ptr(), n, idx.data_ptr(), out.data_ptr(), _st()))
    torch.cuda.synchronize()
    assert int(idx[0]) == hi_idx
    x[0] = float("nan")
    trn.check(L.trn_argmax_f32_dev(x.data_ptr(), n, idx.data_ptr(), out.data_ptr(), _st()))
    torch.cuda.synchronize()
    assert int(idx[0]) == 0


# ---- config 5: softmax / log_softmax / gelu over 4096 x 32000, matmul 32768^3 row block --------------------
def test_config5_row_kernels_4096x32000(trn, gpu, oracle):
    from oracle import SCALAR
    rows, cols = 4096, 32000
    L = trn.lib
    g = torch.Generator(device=gpu); g.manual_seed(0x5EED0007)
    x = torch.randn(rows, cols, device=gpu, generator=g) * 4
    y = torch.empty_like(x)
    trn.check(L.trn_softmax_rows_f32_dev(x.data_ptr(), y.data_ptr(), rows, cols, _st()))
    torch.cuda.synchronize()
    assert (y.double().sum(1) - 1).abs().max().item() < 1e-5            # rows sum to 1
    assert torch.equal(y.argmax(1), x.argmax(1))                        # order preserved
    assert (y >= 0).all() and (y <= 1).all()
    arg = (x - x.max(1, keepdim=True).values).double()
    truth = torch.softmax(arg, dim=1)
    assert (y.double() - truth).abs().max().item() <= 1e-6
    # sampled rows against the scalar-libm oracle (reference semantics, src/vector.rs:1540-1553)
    pick = [0, 1, 147, 148, 2047, 4095]
    want = oracle.softmax_rows(x[pick].cpu().numpy(), len(pick), cols, backend=SCALAR)
    got = y[pick].cpu().numpy()
    assert np.all(np.abs(got - want) <= 1e-6 + 2e-4 * want)
    # log_softmax == log(softmax) and exp(log_softmax) sums to 1
    z = torch.empty_like(x)
    trn.check(L.trn_log_softmax_rows_f32_dev(x.data_ptr(), z.data_ptr(), rows, cols, _st()))
    torch.cuda.synchronize()
    tlog = torch.log_softmax(arg, dim=1)
    assert ((z.double() - tlog).abs() - 4 * torch.finfo(torch.float32).eps * tlog.abs()).max().item() <= 2.0 ** -20
    assert (z.double().exp().sum(1) - 1).abs().max().item() < 1e-5
    del truth, tlog, arg
    # gelu: reference formula in f64 on the device; gelu(0) == 0 exactly
    x[0, 0] = 0.0
    trn.check(L.trn_gelu_f32_dev(x.data_ptr(), x.numel(), y.data_ptr(), _st()))
    torch.cuda.synchronize()
    xd = x.double()
    ref = 0.5 * xd * (1 + torch.tanh(0.7978845608028654 * (xd + 0.044715 * xd ** 3)))
    tol = 4 * torch.finfo(torch.float32).eps * ref.abs() + 4 * 2.0 ** -24 * xd.abs() + 1e-30
    assert ((y.double() - ref).abs() <= tol).all()
    assert float(y[0, 0]) == 0.0
    # sigmoid on the same buffer: monotone, in [0, 1], sigmoid(x) + sigmoid(-x) == 1 within 2 ulp
    trn.check(L.trn_sigmoid_f32_dev(x.data_ptr(), x.numel(), y.data_ptr(), _st()))
    nx = -x
    trn.check(L.trn_sigmoid_f32_dev(nx.data_ptr(), nx.numel(), z.data_ptr(), _st()))
    torch.cuda.synchronize()
    assert ((y.double() + z.double() - 1).abs().max().item()) <= 3e-7
    assert ((y.double() - torch.sigmoid(xd)).abs() <= 4 * torch.finfo(torch.float32).eps * torch.sigmoid(xd)).all()


def test_config5_matmul_32768_row_block(trn, gpu):
    """One GPU's share of the 32768^3 product at 8 GPUs: A-block 4096 x 32768, full B, C-block 4096 x 32768."""
    n, mb = 32768, 4096
    free, _ = torch.cuda.mem_get_info()
    if free < 20 << 30:
        pytest.skip("needs ~20 GiB of free HBM")
    a = u01(0x5EED0008, mb * n, gpu).view(mb, n) * 2 - 1
    b = u01(0x5EED0009, n * n, gpu).view(n, n) * 2 - 1
    c = torch.empty(mb, n, device=gpu)
    trn.check(trn.lib.trn_matmul_f32_dev(a.data_ptr(), mb, n, b.data_ptr(), n, n, c.data_ptr(), _st()))
    torch.cuda.synchronize()
    rows = torch.tensor([0, 127, 128, 2049, 4095], device=gpu)
    cols = torch.arange(0, n, 61, device=gpu)
    bs = b[:, cols].double()
    truth = a[rows].double() @ bs
    scale = a[rows].double().abs() @ bs.abs()
    assert ((c[rows][:, cols].double() - truth).abs() / scale).max().item() <= 1e-5
