"""CPU: the integer identities the match kernels rest on (uzliti_slam_b200/csrc/uz_knn2.cuh), restated in numpy.
The kernels themselves are checked bit for bit against the oracle on the GPU; this keeps the algebra honest on a box
without one: the CSA layout is an invertible XOR transform, the four bit planes it yields reproduce the Hamming distance,
the packed keys order by (distance, train index), and the column candidate of the fused cross-check orders by
(distance, query index)."""
import numpy as np


def csa_pack(w):
    """raw 8 words -> CSA layout (csa_pack in uz_knn2.cuh)"""
    s1 = w[..., 0] ^ w[..., 1] ^ w[..., 2]
    s2 = w[..., 3] ^ w[..., 4] ^ w[..., 5]
    return np.stack([w[..., 0], w[..., 1], s1, w[..., 3], w[..., 4], s2, s1 ^ s2 ^ w[..., 6], w[..., 7]], -1)


def popc(x):
    return np.array([bin(int(v)).count("1") for v in np.asarray(x).ravel()], np.int64).reshape(np.shape(x))


def maj_implied(a, b, s):
    """lop3 0xD4: maj(a, b, c) with the third operand implied by s = a ^ b ^ c"""
    c = a ^ b ^ s
    return (a & b) | (a & c) | (b & c)


def csa_distance(U, T):
    """distance of two CSA-layout rows as csa_key / csa_acc16 compute it: 13 logic ops + 4 popcounts"""
    X = U ^ T
    x0, x1, S1, x3, x4, S2, S3, x7 = (X[..., i] for i in range(8))
    C1 = maj_implied(x0, x1, S1)
    C2 = maj_implied(x3, x4, S2)
    C3 = maj_implied(S1, S2, S3)
    S5 = C1 ^ C2 ^ C3
    C5 = (C1 & C2) | (C1 & C3) | (C2 & C3)
    return popc(S3) + popc(x7) + 2 * popc(S5) + 4 * popc(C5)


def test_csa_planes_reproduce_the_hamming_distance():
    rng = np.random.default_rng(0)
    q = rng.integers(0, 2 ** 32, (300, 8), dtype=np.uint64).astype(np.uint32)
    t = rng.integers(0, 2 ** 32, (300, 8), dtype=np.uint64).astype(np.uint32)
    q[:20, 1:] = 0; t[:20, 1:] = 0; t[5] = ~q[5]; t[6] = q[6]          # low entropy, complement (256), equal rows (0)
    want = popc(q ^ t).sum(-1)
    got = csa_distance(csa_pack(q), csa_pack(t))
    assert np.array_equal(got, want) and want[5] == 256 and want[6] == 0
    # 512-bit rows: two independently transformed halves feed one sum
    q2 = np.concatenate([q, q[::-1]], -1); t2 = np.concatenate([t, t[::-1] ^ np.uint32(0x0F0F0F0F)], -1)
    got2 = csa_distance(csa_pack(q2[:, :8]), csa_pack(t2[:, :8])) + csa_distance(csa_pack(q2[:, 8:]), csa_pack(t2[:, 8:]))
    assert np.array_equal(got2, popc(q2 ^ t2).sum(-1))


def test_csa_layout_is_invertible():
    rng = np.random.default_rng(1)
    w = rng.integers(0, 2 ** 32, (100, 8), dtype=np.uint64).astype(np.uint32)
    c = csa_pack(w)
    back = c.copy()
    back[:, 2] = c[:, 2] ^ c[:, 0] ^ c[:, 1]
    back[:, 5] = c[:, 5] ^ c[:, 3] ^ c[:, 4]
    back[:, 6] = c[:, 6] ^ c[:, 2] ^ c[:, 5]
    assert np.array_equal(back, w)


def test_packed_keys_order_like_opencv():
    """key16 = (distance << shift) | (row & mask) inside a key block, widened to (distance << 16) | trainIdx: the minimum
    of keys is the minimum by (distance, trainIdx), 256-bit rows (7-bit rows, distance <= 256) and 512-bit rows (6-bit
    rows, distance <= 512); both fit 16 bits with 0xFFFF left as "none"."""
    for shift, dmax in ((7, 256), (6, 512)):
        mask = (1 << shift) - 1
        assert (dmax << shift) | mask < 0xFFFF
        rng = np.random.default_rng(shift)
        d = rng.integers(0, dmax + 1, 4000)
        row = rng.integers(0, 4096, 4000)
        base = row & ~mask
        key16 = (d << shift) | (row & mask)
        wide = ((key16 >> shift) << 16) | (base + (key16 & mask))          # merge_block16 / merge_block16w
        assert np.array_equal(wide, (d << 16) | row)
        order = np.lexsort((row, d))
        assert np.array_equal(np.argsort(wide, kind="stable"), order) or np.array_equal(np.sort(wide), wide[order])


def test_cross_check_column_candidate_orders_by_distance_then_query():
    """col_update16: v = (key16 << 16) | local query index; both keys of a thread carry the same train row, so the minimum
    over queries is the minimum by (distance, query index); col_flush recovers (distance << 16) | queryIdx."""
    for shift in (7, 6):
        rng = np.random.default_rng(10 + shift)
        row = 37 & ((1 << shift) - 1)
        d = rng.integers(0, 40, 512)                                       # many ties
        idx = np.arange(512)
        v = (((d << shift) | row) << 16) | idx
        best = int(v.min())
        want = min(zip(d.tolist(), idx.tolist()))
        q_base = 1024
        flushed = ((best >> (16 + shift)) << 16) | (q_base + (best & 0xFFFF))
        assert flushed == (want[0] << 16) | (q_base + want[1])
