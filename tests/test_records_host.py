"""Host-side decoders of the packed step records (no GPU): records are built here bit by bit from the layouts stated in
include/dmp.h and must decode to the rows they were built from."""
import numpy as np
import pytest

from snac_b200 import _lib as L
from snac_b200.vecenv import bits_bytes, record_dtype, unpack_bits, unpack_records


def encode_bits(dim, obs, reward, done, saturated):
    """dmp.h: field i of width b at bits [i b, i b + b) of a little-endian bit string; 2D trailer at bit 98, 3D at bit 224."""
    n = len(obs)
    cw = 2 if dim == 2 else 4
    out = np.zeros((n, bits_bytes(dim)), np.uint8)
    for e in range(n):
        v = 0
        for j in range(49):
            v |= int(obs[e, j] + 1) << (j * cw)
        tr = int(obs[e, 49]) | int(obs[e, 50]) << 12 | L.BITS_REWARDS.index(float(reward[e])) << 24 | int(done[e]) << 27 | int(saturated[e]) << 28
        v |= tr << (98 if dim == 2 else 224)
        out[e] = np.frombuffer(v.to_bytes(bits_bytes(dim), "little"), np.uint8)
    return out


@pytest.mark.parametrize("dim", [2, 3])
def test_unpack_bits_inverts_the_header_layout(dim):
    rng = np.random.RandomState(dim)
    n = 257
    obs = np.zeros((n, 51))
    obs[:, :49] = rng.randint(-1, 2 if dim == 2 else 15, size=(n, 49))
    obs[:, 49], obs[:, 50] = rng.randint(0, 4096, size=n), rng.randint(0, 4096, size=n)
    reward = np.asarray(L.BITS_REWARDS[:6], np.float32)[rng.randint(0, 6, size=n)]
    done, sat = rng.randint(0, 2, size=n).astype(bool), rng.randint(0, 2, size=n).astype(bool)
    rec = encode_bits(dim, obs, reward, done, sat)
    assert rec.shape == (n, 16 if dim == 2 else 32)
    o, r, d, s = unpack_bits(rec, dim)
    assert np.array_equal(o, obs) and np.array_equal(r, reward) and np.array_equal(d, done) and np.array_equal(s, sat)
    o2, r2, d2, s2 = unpack_records(rec.reshape(1, n, -1), dim, np.float32)          # dispatch by record size, leading axes kept
    assert o2.shape == (1, n, 51) and o2.dtype == np.float32 and np.array_equal(o2[0], obs.astype(np.float32))
    assert np.array_equal(r2[0], reward) and np.array_equal(d2[0], done) and np.array_equal(s2[0], sat)


def test_byte_record_fields_sit_where_the_header_says():
    rec = np.zeros((3, 56), np.uint8)
    rec[:, :49] = np.arange(49) % 3
    rec[:, 49] = [0, L.REC_DONE, L.REC_SATURATED]
    rec[:, 50:52] = np.array([[0x34, 0x12]] * 3, np.uint8)      # count_brick = 0x1234, little endian
    rec[:, 52:54] = np.array([[0x02, 0x01]] * 3, np.uint8)      # count_step = 0x0102
    rec[:, 54] = np.array([5, 0xFF, 0x9C], np.uint8)            # rewards 5, -1, -100 as i8
    rec[:, 55] = [0, 1, 0]
    assert record_dtype(2).itemsize == 56 and record_dtype(1).itemsize == 16
    o, r, d, s = unpack_records(rec, 2)
    assert np.array_equal(o[:, :49], (np.arange(49) % 3 - 1.0) * np.ones((3, 1))) and (o[:, 49] == 0x1234).all() and (o[:, 50] == 0x0102).all()
    assert np.array_equal(r, [5, -1, -100]) and np.array_equal(d, [False, True, False]) and np.array_equal(s, [False, False, True])
    rec1 = np.zeros((2, 16), np.uint8)
    rec1.view("<i2")[:, :5] = [[-1, -1, 0, 3, 7], [2, 2, 2, -1, -1]]
    rec1.view("<u2")[:, 5], rec1.view("<u2")[:, 6] = [9, 10], [700, 701]
    rec1[:, 14], rec1[:, 15] = [10, 0], [1, 0]
    o, r, d, s = unpack_records(rec1, 1)
    assert np.array_equal(o, [[-1, -1, 0, 3, 7, 9, 700], [2, 2, 2, -1, -1, 10, 701]]) and np.array_equal(r, [10, 0]) and np.array_equal(d, [True, False])


def test_bit_records_exist_for_2d_and_3d_only():
    with pytest.raises(ValueError):
        bits_bytes(1)
    with pytest.raises(ValueError):
        unpack_bits(np.zeros((4, 56), np.uint8), 2)
