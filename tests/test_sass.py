"""CPU: the shipped library's SASS carries the instruction mix the roofline arithmetic assumes (uzliti_slam_b200/mix.py), and
the match kernels are what DESIGN.md says they are: TMA bulk copies + mbarriers in both, POPC/LOP3 on the integer pipes in
knn2_kernel, tcgen05.mma (UTCIMMA) + TMEM loads (LDTM) and NO popcount in knn2_mma_kernel."""
import re
import shutil
import subprocess

import pytest

from uzliti_slam_b200 import binding, mix


def _functions(built):
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not on PATH")
    sass = subprocess.check_output(["cuobjdump", "-sass", binding.lib_path()], text=True)
    funcs, name, body = {}, None, []
    for ln in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            if name:
                funcs[name] = body
            name, body = m.group(1), []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
        if m and name:
            body.append((int(m.group(1), 16), m.group(2).strip()))
    if name:
        funcs[name] = body
    return funcs


def _op(text):
    t = text.split()
    return t[1] if t[0].startswith("@") else t[0]


def _hot_loop(body):
    """the main compare loop: the smallest body of a backward branch with at least 16 POPC and packed (U16x2) min/max"""
    addr_index = {a: i for i, (a, _) in enumerate(body)}
    best = None
    for i, (a, text) in enumerate(body):
        m = re.search(r"\bBRA(?:\.\w+)*\s+(?:!?\w+,\s*)?(0x[0-9a-f]+)", text)
        if not m:
            continue
        tgt = int(m.group(1), 16)
        if tgt < a and tgt in addr_index:
            j = addr_index[tgt]
            n = sum(1 for _, t in body[j:i + 1] if _op(t).startswith("POPC"))
            packed = any("U16x2" in _op(t) for _, t in body[j:i + 1])
            size = i - j
            if n >= 16 and packed and (best is None or size < best[1]):
                best = (n, size, j, i)
    return best


def test_integer_pipe_kernel_mix_matches_the_roofline_constants(built):
    funcs = _functions(built)
    # knn2_kernel<256, 2, true, true, false, false>
    names = [n for n in funcs if "knn2_kernel" in n and "ILi256ELi2ELb1ELb1ELb0ELb0E" in n]
    assert len(names) == 1, [n for n in funcs if "knn2_kernel" in n][:4]
    body = funcs[names[0]]
    hot = _hot_loop(body)
    assert hot is not None
    _, _, j, i = hot
    ops = [_op(t) for _, t in body[j:i + 1]]
    cnt = lambda p: sum(1 for o in ops if o.startswith(p))          # noqa: E731
    n = mix.KNN2_COMPARES_PER_ITERATION
    assert cnt("POPC") == mix.KNN2_POPC * n
    assert cnt("LOP3") == mix.KNN2_LOP3 * n
    assert mix.KNN2_IMAD * n <= cnt("IMAD") <= mix.KNN2_IMAD * n + 2
    assert cnt("VIMNMX") == int(mix.KNN2_MINMAX * n)
    assert cnt("LDS") == 8 and all("128" in t for _, t in body[j:i + 1] if _op(t).startswith("LDS"))
    whole = [_op(t) for _, t in body]
    assert any(o.startswith("UBLKCP") for o in whole) and any(o.startswith("SYNCS") for o in whole)      # TMA bulk copy + mbarrier


def test_default_match_kernel_is_block_scaled_fp4_tcgen05(built):
    """knn2_mmaf_kernel: tcgen05.mma kind::mxf4 (UTCOMMA) started by one kind::f8f6f4 instruction (UTCQMMA), issued back to
    back by an elected lane, packed TMEM loads, an epilogue of three-input packed max and one multiply-add per register"""
    funcs = _functions(built)
    names = [n for n in funcs if "knn2_mmaf_kernelILb0E" in n]          # <false>: 32-byte rows
    assert len(names) == 1
    wide = [n for n in funcs if "knn2_mmaf_kernelILb1E" in n]           # <true>: 64-byte rows, eight instructions per accumulator
    assert len(wide) == 1
    wops = [_op(t) for _, t in funcs[wide[0]]]
    assert sum(1 for o in wops if o.startswith("UTCOMMA")) == 2 * sum(1 for o in wops if o.startswith("UTCQMMA")) * mix.F4_INSTRUCTIONS_PER_TILE
    body = funcs[names[0]]
    ops = [_op(t) for _, t in body]
    cnt = lambda p: sum(1 for o in ops if o.startswith(p))          # noqa: E731
    assert cnt("UTCOMMA") >= mix.F4_INSTRUCTIONS_PER_TILE and cnt("UTCOMMA") % mix.F4_INSTRUCTIONS_PER_TILE == 0
    assert cnt("UTCQMMA") * mix.F4_INSTRUCTIONS_PER_TILE == cnt("UTCOMMA") * mix.F4_START_INSTRUCTIONS
    assert cnt("UTCIMMA") == 0 and cnt("HMMA") == 0 and cnt("POPC") <= 2
    assert sum(1 for o in ops if o.startswith("LDTM") and "PACK16BIT" in o) >= 2 and cnt("UBLKCP") >= 3 and cnt("UTCBAR") >= 3
    # the instructions of one accumulator come out back to back: nothing but uniform-datapath moves between them
    idx = [k for k, o in enumerate(ops) if o.startswith("UTCOMMA")]
    for a, b in zip(idx[:-1], idx[1:]):
        if b - a < 40:                                    # same accumulator
            assert all(o.startswith(("UMOV", "UIADD3", "ULOP3", "USHF", "ULEA", "USEL", "NOP")) for o in ops[a + 1:b]), ops[a + 1:b]
    # per accumulator half of 128 columns (64 registers): 2 x 32 three-input packed max and 64 multiply-adds
    max3 = sum(1 for o in ops if o.startswith("VIMNMX3") and "U16x2" in o)
    assert max3 >= int((64 + 56) * 2 * 2 * mix.F4_EPILOGUE_MINMAX3)
    assert cnt("IMAD") >= int((64 + 56) * 2 * mix.F4_EPILOGUE_IMAD)


def test_int8_tensor_core_kernel_is_tcgen05_with_tmem_and_no_popcount(built):
    funcs = _functions(built)
    names = [n for n in funcs if "knn2_mmak_kernel" in n]
    assert len(names) == 1
    body = funcs[names[0]]
    ops = [_op(t) for _, t in body]
    cnt = lambda p: sum(1 for o in ops if o.startswith(p))          # noqa: E731
    per_step = mix.MMA_INSTRUCTIONS_PER_TILE + mix.MMA_KEY_SLICE_INSTRUCTIONS
    assert cnt("UTCIMMA") >= per_step and cnt("UTCIMMA") % per_step == 0          # 8 x K = 32 of the descriptor + the constant key slice
    assert cnt("LDTM") >= 4 and sum(1 for o in ops if o.startswith("LDTM") and "PACK16BIT" in o) >= 2   # packed loads: two columns per register
    assert cnt("UBLKCP") >= 3 and cnt("SYNCS") >= 10 and cnt("UTCBAR") + cnt("UTCCOMMIT") + cnt("UTC") >= 1
    assert cnt("POPC") <= 2 and cnt("HMMA") == 0      # (a stray POPC of the warp-vote lowering; the compares use none)
    # the epilogue is packed max only: per 128 columns of the full path 64 registers x 2.5 packed min/max, and NO multiply-add
    # that builds a key (the IMAD epilogue of knn2_mma_kernel carries immediates 0x40xx40xx = 16384 + column pair)
    packed = sum(1 for o in ops if o.startswith("VIMNMX") and "U16x2" in o)
    assert packed >= int(64 * 2 * mix.MMA_EPILOGUE_MINMAX)
    assert not [t for _, t in body if _op(t).startswith("IMAD") and re.search(r"0x40[0-9a-f]{2}40[0-9a-f]{2}\b", t)]
    # the measured alternative still has them
    alt = [n for n in funcs if "knn2_mma_kernel" in n]
    assert len(alt) == 1
    assert len([t for _, t in funcs[alt[0]] if _op(t).startswith("IMAD") and re.search(r"0x40[0-9a-f]{2}40[0-9a-f]{2}\b", t)]) >= 64 * 2
