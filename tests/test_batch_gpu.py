"""CommandBatch (trn_batch_*): the CUDA counterpart of GpuCommandBatch, checked with the reference's own batch
tests (src/backends/gpu/batch.rs:1046-1240) and bit-for-bit against the one-op-at-a-time `_dev` path."""
import numpy as np
import pytest

from oracle import SCALAR

pytestmark = pytest.mark.gpu
f32 = np.float32


def test_buffer_allocation_and_operation_queuing(trn):
    b = trn.CommandBatch()                                   # batch.rs:1046-1078
    b1, b2 = b.upload([1, 2, 3]), b.upload([4, 5, 6])
    assert b.num_buffers() == 2 and b1 != b2
    b = trn.CommandBatch()
    x = b.upload([1, 2, -3, 4])
    r = b.relu(x)
    s = b.scale(r, 2.0)
    o = b.upload([0.5] * 4)
    b.add(s, o)
    assert b.num_operations() == 3 and b.num_buffers() == 5


@pytest.mark.parametrize("op", ["add", "mul", "dot", "sub"])
def test_size_mismatch(trn, op):
    b = trn.CommandBatch()                                   # batch.rs:1080-1117 (the reference panics)
    a, c = b.upload([1, 2]), b.upload([1, 2, 3])
    with pytest.raises(trn.TruenoError) as e:
        getattr(b, op)(a, c)
    assert e.value == trn.TruenoError.InvalidInput("Buffer size mismatch: 2 vs 3")


def test_all_batch_operations(trn):
    b = trn.CommandBatch()                                   # batch.rs:1120-1240
    add_result = b.add(b.scale(b.relu(b.upload([1, 2, -3, 4])), 2.0), b.upload([0.5] * 4))
    mul_result = b.mul(b.upload([1, 2, 3, 4]), b.upload([2, 3, 4, 5]))
    dot_result = b.dot(b.upload([1, 2, 3, 4]), b.upload([2, 3, 4, 5]))
    sig_result = b.sigmoid(b.upload([-2, -1, 0, 1, 2]))
    tanh_result = b.tanh(b.upload([-1, 0, 1]))
    swish_result = b.swish(b.upload([0, 1, 2]))
    gelu_result = b.gelu(b.upload([-1, 0, 1]))
    sub_result = b.sub(b.upload([5, 10, 15, 20]), b.upload([1, 2, 3, 4]))
    chain_result = b.tanh(b.sigmoid(b.relu(b.upload([-2, -1, 0, 1, 2]))))
    with pytest.raises(trn.TruenoError):
        b.read(add_result)                                   # not executed yet
    launches0 = trn.launch_count()
    b.execute()
    assert trn.launch_count() - launches0 == b.num_operations() == 13
    assert np.allclose(b.read(add_result), [2.5, 4.5, 0.5, 8.5], atol=1e-5)
    assert b.read(mul_result).tolist() == [2, 6, 12, 20]
    assert b.read(dot_result).tolist() == [40.0]
    r = b.read(sig_result)
    assert abs(r[0] - 0.119) < 0.01 and abs(r[2] - 0.5) < 0.01 and abs(r[4] - 0.881) < 0.01
    r = b.read(tanh_result)
    assert abs(r[0] + 0.762) < 0.01 and abs(r[1]) < 0.01 and abs(r[2] - 0.762) < 0.01
    r = b.read(swish_result)
    assert abs(r[0]) < 0.01 and abs(r[1] - 0.731) < 0.01
    r = b.read(gelu_result)
    assert abs(r[1]) < 0.01 and abs(r[2] - 0.841) < 0.05
    assert b.read(sub_result).tolist() == [4, 8, 12, 16]
    r = b.read(chain_result)
    assert r.size == 5 and ((r >= 0) & (r <= 1)).all()


def test_batch_matches_single_ops_and_replays(trn, oracle):
    """A long chain on 1M elements: bit-identical to the host-slice ops run one by one, and the instantiated graph
    replays on new inputs after update()."""
    n = (1 << 20) + 5
    rng = np.random.default_rng(3)
    x, y = rng.standard_normal(n).astype(f32), rng.standard_normal(n).astype(f32)

    def build(b):
        ix, iy = b.upload(x), b.upload(y)
        h = b.gelu(b.add(b.mul(ix, iy), ix))
        return ix, iy, b.sub(b.swish(h), b.scale(b.tanh(iy), 0.5)), b.dot(h, iy)

    def expect(xv, yv):
        V = trn.Vector
        h = V(xv).mul(V(yv)).add(V(xv)).gelu()
        out = h.swish().sub(V(yv).tanh().scale(0.5))
        return out.as_slice(), h.dot(V(yv))

    b = trn.CommandBatch()
    ix, iy, out_id, dot_id = build(b)
    b.execute()
    want_out, want_dot = expect(x, y)
    assert np.array_equal(b.read(out_id), want_out)
    assert b.read(dot_id)[0] == want_dot
    # against the oracle too (gelu tolerance of tests/test_parity_gpu.py)
    g = oracle.gelu((x * y + x).astype(f32), backend=SCALAR)
    h = trn.Vector(x).mul(trn.Vector(y)).add(trn.Vector(x)).gelu().as_slice()
    assert np.max(np.abs(h - g)) <= 1e-5
    # replay on new data: one graph launch, no re-capture
    x2 = rng.standard_normal(n).astype(f32)
    b.update(ix, x2)
    launches0 = trn.launch_count()
    b.execute()
    assert trn.launch_count() - launches0 == b.num_operations()
    want_out2, want_dot2 = expect(x2, y)
    assert np.array_equal(b.read(out_id), want_out2) and b.read(dot_id)[0] == want_dot2
