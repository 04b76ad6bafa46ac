"""Host-side restatement of the arithmetic knn2_mmaf_kernel (uz_knn2_mmaf.cuh) relies on, checked exhaustively or on random
data with numpy - no GPU.  The kernel's bit-exact parity with the oracle is tests/test_mma.py (-m gpu); these tests pin WHY it
is exact, so that a change of a constant shows up here first.
  * the fp32 accumulator: 2^23 + 16384 + 127 - column + 64 * dot is exact in float32 whatever the order of the partial sums,
    and its low 16 bits are 32895 - ((hamming << 7) | column);
  * the start operands: 127 - (row & 127) = 64 d2 + 8 d1 + d0 with digits 0..7, all operands exact in e5m2;
  * the epilogue's second sweep: y = x - P + 65536 as ONE 32-bit add on registers holding two 16-bit keys finds the runner-up
    of both lanes, given that high-lane keys are even (odd columns, odd offset);
  * a ragged block: keys 0 in the missing columns never displace a real key;
  * the E4 operand layout is a bijection onto e4_bytes(n) and 8-row groups are contiguous."""
import numpy as np

TOP = 32895


def test_accumulator_is_exact_in_float32_and_carries_the_key():
    ham = np.arange(0, 257, dtype=np.int64)[:, None]
    col = np.arange(0, 128, dtype=np.int64)[None, :]
    dot = 256 - 2 * ham                                   # <q, t> of +-1 vectors
    start = np.float32(2.0 ** 23) + np.float32(16384.0) + (np.float32(127.0) - col.astype(np.float32))
    # four partial sums of 64 products each, any split of the dot product: every partial value is an integer below 2^24
    rng = np.random.default_rng(1)
    for _ in range(8):
        parts = rng.multinomial(256, [0.25] * 4)          # how many bits each K = 64 instruction covers is fixed (64); the
        acc = start.copy()                                # split of AGREEMENTS among them is what varies
        left = dot.copy()
        for k in range(4):
            lo = np.maximum(-64, left - 64 * (3 - k))
            hi = np.minimum(64, left + 64 * (3 - k))
            part = np.clip(left * parts[k] // 256, lo, hi) if k < 3 else left
            part = np.clip(part, lo, hi)
            acc = (acc + np.float32(64.0) * part.astype(np.float32)).astype(np.float32)
            left = left - part
        assert (left == 0).all()
        bits = acc.view(np.uint32)
        assert ((bits >> 23) == 150).all()                # exponent of [2^23, 2^24): the mantissa is the integer
        key16 = (ham << 7) | col
        assert ((bits & 0xFFFF).astype(np.int64) == TOP - key16).all()
    # the complement orders like OpenCV: smaller distance first, then smaller column
    key16 = ((ham << 7) | col).ravel()
    assert (np.argsort(-(TOP - key16), kind="stable") == np.argsort(key16, kind="stable")).all()


def _e5m2(v):
    """value of an e5m2 byte (no NaN / inf patterns used here)"""
    s, e, m = v >> 7, (v >> 2) & 31, v & 3
    mag = (m / 4.0) * 2.0 ** -14 if e == 0 else (1 + m / 4.0) * 2.0 ** (e - 15)
    return -mag if s else mag


def test_start_operands_are_exact_e5m2():
    dig = [0x00, 0x3C, 0x40, 0x42, 0x44, 0x45, 0x46, 0x47]
    assert [_e5m2(b) for b in dig] == [0, 1, 2, 3, 4, 5, 6, 7]
    assert _e5m2(0x68) == 2048 and _e5m2(0x58) == 128 and _e5m2(0x54) == 64 and _e5m2(0x48) == 8 and _e5m2(0x3C) == 1
    a_tail = [2048, 2048, 128, 64, 8, 1]
    for row in range(240):
        v = 127 - (row & 127)
        b_tail = [2048, 2048, 128, v >> 6, (v >> 3) & 7, v & 7]
        assert max(b_tail[3:]) <= 7
        assert sum(x * y for x, y in zip(a_tail, b_tail)) == 2 ** 23 + 16384 + 127 - (row & 127)
    # a missing column of a ragged tile starts - and, with zero operand rows, ends - at exactly 2^23: key 0
    assert sum(x * y for x, y in zip(a_tail, [2048, 2048, 0, 0, 0, 0])) == 2 ** 23


def _second_sweep(regs):
    """the kernel's two sweeps on uint32 registers [rows, R] holding (odd-column key << 16) | even-column key"""
    lo, hi = regs & 0xFFFF, regs >> 16
    P = (hi.max(axis=1) << 16) | lo.max(axis=1)
    c = (65536 - P) & 0xFFFFFFFF
    y = (regs + c[:, None]) & 0xFFFFFFFF
    Y = ((y >> 16).max(axis=1) << 16) | (y & 0xFFFF).max(axis=1)
    S = ((Y - c) & 0xFFFFFFFF) & 0xFFFEFFFF
    return P, S


def _true_top2(regs):
    lo, hi = np.sort(regs & 0xFFFF, axis=1), np.sort(regs >> 16, axis=1)
    return (hi[:, -1] << 16) | lo[:, -1], (hi[:, -2] << 16) | lo[:, -2]


def _random_block(rng, rows, R, n_real):
    """keys of a 128-column block: distinct per lane, even in the high lane, 0 in missing columns"""
    ham = rng.integers(0, 257, (rows, 2 * R))
    if rng.integers(0, 2):
        ham = rng.integers(100, 104, (rows, 2 * R))       # tie-heavy
    col = np.arange(2 * R)[None, :]
    key = TOP - ((ham << 7) | col)
    key[:, n_real:] = 0
    return (key[:, 1::2].astype(np.uint64) << 16 | key[:, 0::2].astype(np.uint64)).astype(np.int64)


def test_one_32_bit_add_finds_the_runner_up_of_both_lanes():
    rng = np.random.default_rng(7)
    for R, n_real in [(64, 128), (56, 112), (64, 127), (64, 40), (64, 5), (64, 4), (56, 17), (64, 65)]:
        for _ in range(20):
            regs = _random_block(rng, 512, R, n_real)
            assert ((regs >> 16) % 2 == 0).all()          # odd columns: 32895 - (.. | odd) is even
            P, S = _second_sweep(regs)
            tp, ts = _true_top2(regs)
            assert (P == tp).all()
            assert (S == ts).all(), (R, n_real)
    # the register that holds the low lane's winner does not borrow: force it to hold the high lane's runner-up as well
    regs = _random_block(rng, 256, 64, 128)
    lo, hi = regs & 0xFFFF, regs >> 16
    j = lo.argmax(axis=1)
    second = np.sort(hi, axis=1)[:, -2]
    k = (hi == second[:, None]).argmax(axis=1)
    rows = np.arange(256)
    hi[rows, k], hi[rows, j] = hi[rows, j].copy(), hi[rows, k].copy()
    regs = (hi << 16) | lo
    P, S = _second_sweep(regs)
    tp, ts = _true_top2(regs)
    assert (P == tp).all() and (S == ts).all()


def test_missing_columns_never_displace_a_real_key():
    rng = np.random.default_rng(9)
    for n_real in range(4, 128):
        regs = _random_block(rng, 64, 64, n_real)
        real = _random_block(rng, 64, 64, 128)           # (shape only)
        P, S = _second_sweep(regs)
        lo, hi = regs & 0xFFFF, regs >> 16
        n_lo, n_hi = (n_real + 1) // 2, n_real // 2
        assert (np.sort(lo[:, :n_lo], axis=1)[:, -2] == (S & 0xFFFF)).all()
        assert (np.sort(hi[:, :n_hi], axis=1)[:, -2] == (S >> 16)).all()
        assert ((P & 0xFFFF) > 0).all() and ((S & 0xFFFF) > 0).all() and ((P >> 16) > 0).all() and ((S >> 16) > 0).all()
        del real
    # a real key of a ragged block is never 0: 0 would be distance 256 in column 127, and a ragged block ends before it
    assert TOP - ((256 << 7) | 126) == 1 and TOP - ((256 << 7) | 111) == 16


def _e4_offset(i, k):
    return (i >> 3) * 1024 + (k >> 5) * 128 + (i & 7) * 16 + ((k >> 1) & 15)


def test_e4_layout_is_a_bijection_with_contiguous_row_groups():
    n = 77
    seen = set()
    for i in range(n):
        for k in range(0, 256, 2):                        # two 4-bit values per byte
            o = _e4_offset(i, k)
            assert _e4_offset(i, k + 1) == o and o not in seen
            seen.add(o)
    groups = (n + 7) // 8
    assert max(seen) < groups * 1024                      # e4_bytes(n)
    for g in range(n // 8):                               # a full 8-row group fills its 1 KB exactly: one bulk copy per run of groups
        assert {o for o in seen if g * 1024 <= o < (g + 1) * 1024} == set(range(g * 1024, (g + 1) * 1024))


def _ring_violations(T, stages, bufs, wait_rule, items=40):
    """A model of the producer of knn2_mmaf_kernel: train tile v may be filled once tile v - stages is done (the stage barrier);
    the query tiles of item k go into buffer k % bufs and may be filled once the LAST tile of item k - bufs is done.  The
    producer only knows what its own waits told it.  `wait_rule(uB, u)` says whether it waits on the stage barrier of tile u
    before loading query tiles when uB tiles have been filled.  Returns the loads that happened without that knowledge."""
    known_done = -1                       # the producer knows tiles <= known_done are done
    uB, bad = 0, 0
    last_of = {}
    for k in range(items):
        b = k % bufs
        if b in last_of:
            u = last_of[b]
            if wait_rule(uB, u):
                known_done = max(known_done, u)
            if known_done < u:
                bad += 1
        for _ in range(T):
            if uB - stages >= 0:
                known_done = max(known_done, uB - stages)      # the wait before refilling the stage
            uB += 1
        last_of[b] = uB - 1
    return bad


def test_query_tiles_are_released_by_the_stage_barrier_of_the_last_train_tile():
    rule = lambda uB, u, S: uB <= u + S                  # noqa: E731  (uz_knn2_mmaf.cuh, producer)
    for stages, bufs in ((4, 2), (2, 1)):
        for T in range(1, 12):
            assert _ring_violations(T, stages, bufs, lambda uB, u: rule(uB, u, stages)) == 0, (T, stages, bufs)
    # the rule never waits on a stage that has been refilled since (its barrier is phases ahead: the parity test would alias)
    for stages, bufs in ((4, 2), (2, 1)):
        for T in range(1, 12):
            uB = 0
            last_of = {}
            for k in range(30):
                b = k % bufs
                if b in last_of and rule(uB, last_of[b], stages):
                    assert uB <= last_of[b] + stages          # tile last + stages, the refill, is not behind us
                uB += T
                last_of[b] = uB - 1
    # the first version skipped the wait when the refill was the very next tile: items of three tiles, four stages, two buffers
    assert _ring_violations(3, 4, 2, lambda uB, u: uB < u + 4) > 0
