"""Pin the noise twin: Random123's published Philox4x32-10 known-answer vectors
(kat_vectors in the Random123 distribution), then the derived mask/normal streams."""
import numpy as np

from oracle import philox


def _p(ctr, key):
    return [int(x) for x in philox.philox4x32_10(*[np.uint32(c) for c in ctr], key[0], key[1])]


def test_philox_known_answers():
    assert _p((0, 0, 0, 0), (0, 0)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    assert _p((f, f, f, f), (f, f)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert _p((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0)) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_mask_is_keyed_and_balanced():
    m = philox.dropout_mask(1234, 3, 7, 5, np.arange(64), 16384)
    assert m.shape == (64, 16384) and set(np.unique(m)) == {0.0, 2.0}
    assert abs(m.mean() - 1.0) < 0.01
    # prefix property: a shorter draw at the same site is a prefix (bits are addressed, not streamed)
    assert np.array_equal(philox.dropout_mask(1234, 3, 7, 5, np.arange(64), 256), m[:, :256])
    # rows are addressed too (sample sharding / row subsets must agree)
    assert np.array_equal(philox.dropout_mask(1234, 3, 7, 5, np.array([9, 40]), 512), m[[9, 40], :512])
    for other in (philox.dropout_mask(1235, 3, 7, 5, np.arange(4), 512),
                  philox.dropout_mask(1234, 4, 7, 5, np.arange(4), 512),
                  philox.dropout_mask(1234, 3, 8, 5, np.arange(4), 512),
                  philox.dropout_mask(1234, 3, 7, 6, np.arange(4), 512)):
        assert not np.array_equal(other, m[:4, :512])


def test_normals_moments():
    z = philox.normals(99, 0, 0, 3, np.arange(20000), 10)
    assert z.dtype == np.float32 and np.isfinite(z).all()
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1.0) < 0.01
    u = np.array([philox.uniform24(5, 0, 0, 40, r) for r in range(2000)])
    assert u.min() >= 0.0 and u.max() < 1.0 and abs(u.mean() - 0.5) < 0.03
