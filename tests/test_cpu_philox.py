"""Pins the host restatement of Philox4x32-10 (tests/support.py) that tests/test_gpu_philox.py compares the kernels with:
known-answer vectors of the published algorithm (Salmon, Moraes, Dror & Shaw, SC'11; Random123 kat_vectors, philox4x32 10)."""
import numpy as np

from tests import support as S


def test_philox4x32_10_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = S.philox4x32_10(*ctr, *key)
        assert tuple(int(v) for v in got) == want


def test_philox_is_vectorised_consistently():
    c0 = np.arange(1000, dtype=np.uint64) * 7919
    many = S.philox4x32_10(c0, 3, 5, 0, 44, 2)
    for i in (0, 1, 17, 999):
        one = S.philox4x32_10(int(c0[i]), 3, 5, 0, 44, 2)
        assert [int(w[i]) for w in many] == [int(w) for w in one]


def test_u01_has_53_bits_and_stays_below_one():
    w0 = np.array([0xffffffff, 0, 0x80000000], np.uint32)
    w1 = np.array([0xffffffff, 0, 0], np.uint32)
    u = S.philox_u01(w0, w1)
    assert u[0] < 1.0 and u[0] == 1.0 - 2.0 ** -53 and u[1] == 0.0 and u[2] == 0.5


def test_streams_for_layout_small_cells():
    """slot e of cell c draws word e % 4 of block e // 4; the pair starting at slot 2k takes words (2k) % 4, +1 of block k // 2 | 2^31"""
    ijk = np.array([0, 0, 0, 0, 0, 0, 2, 2, 2, 5], np.uint32)           # cells of 6, 3 and 1 super-droplets
    sid = np.array([3, 9, 1, 0, 4, 8, 2, 7, 5, 6], np.uint32)
    un, u01 = S.philox_streams_for_layout(sid, ijk, seed=44, call=7, small=True, cell_base=100, stream=3)
    blk = lambda c, q: [int(w) for w in S.philox4x32_10(100 + c, q, 7, 0, 44, 3)]
    assert un[4] == blk(0, 1)[0] and un[8] == blk(0, 1)[1] and un[3] == blk(0, 0)[0] and un[7] == blk(2, 0)[1]
    w = blk(0, 0x80000000)
    assert u01[0] == S.philox_u01(np.uint32(w[0]), np.uint32(w[1])) and u01[2] == S.philox_u01(np.uint32(w[2]), np.uint32(w[3]))
    w = blk(0, 0x80000001)
    assert u01[4] == S.philox_u01(np.uint32(w[0]), np.uint32(w[1]))
    w = blk(2, 0x80000000)
    assert u01[6] == S.philox_u01(np.uint32(w[0]), np.uint32(w[1]))
