"""Host-side invariants of the tensor-core engine's tile tables (ecnf_solve_tc.cuh: tc_pack), no GPU needed.

The kernel relies on them: a tile is 16 (x2 sub-tiles for U = 64) chunks of 8 columns; a chunk is one primal row of a
group plus up to 7 of its tangent rows (slots qb .. qb+cnt-2), or 8 primal rows of consecutive groups (dense kinds); every
(group, slot) row appears exactly once (the primal of a group is repeated in each of its chunks, one of them flagged as
owner); chunk p of a tile lives in sub-tile p % SUB, half (p / SUB) % 2, position p / (2 SUB); a run (the chunks whose
contributions are summed before they touch an accumulator) is a contiguous chunk range of one (receiver, slot group) with
its end and every thread group's last chunk flagged; tiles of the message-passing kinds never span two windows and the
rows of a window fit the accumulator."""
import ctypes as C

import numpy as np
import pytest

from ecnf_b200.engine import CnfConfig, Engine

TILE_WORDS = 48
VALID, GRPEND, OWNER, RUNEND = 1 << 31, 1 << 22, 1 << 23, 1 << 24
N, FLUSH, WIN_I, WIN_NR, WIN_S, WIN_NS, NCH, IFIRST, ILAST, P, MR = 0, 1, 2, 3, 4, 5, 6, 10, 11, 12, 13
NODE1, NODE, FIRST, MID, LAST, EDGE1 = range(6)


def table(eng, kind):
    cnt = eng.lib.ecnf_solve_tc_tile_table(eng.handle, kind, None, 0)
    buf = np.zeros(cnt * TILE_WORDS, np.uint32)
    got = eng.lib.ecnf_solve_tc_tile_table(eng.handle, kind, buf.ctypes.data_as(C.POINTER(C.c_uint32)), buf.size)
    assert got == cnt
    return buf.reshape(cnt, TILE_WORDS)


def word_index(p, sub):
    return (p % sub) * 16 + ((p // sub) % 2) * 8 + p // (2 * sub)


SHAPES = [(13, 3, 128, 64), (4, 2, 128, 64), (2, 2, 128, 64), (7, 3, 128, 64), (22, 3, 64, 32), (5, 3, 64, 32), (9, 2, 64, 32)]


@pytest.mark.parametrize("hutch", [False, True])
@pytest.mark.parametrize("n,dim,U,H", SHAPES)
def test_tile_tables(n, dim, U, H, hutch):
    """hutch: the one-tangent tables of the Hutchinson mode (kinds NODE and MID with two rows per group)."""
    eng = Engine(CnfConfig(n, dim, 0.01, 1.0, 3, (U,) * 2, H, 8, 1))
    SUB = 128 // U
    D, ND, nb = n * dim, (2 if hutch else 1 + n * dim), n - 1
    rs = {NODE1: 1, NODE: ND, FIRST: 1 + 2 * dim, MID: ND, LAST: 1 + dim, EDGE1: 1}
    for kind in ((NODE, MID) if hutch else range(6)):
        tab = table(eng, kind + (8 if hutch else 0))
        assert len(tab) > 0
        r, ngroups = rs[kind], (n if kind < 2 else n * nb)
        seen, owners = set(), set()
        for t in tab:
            h = t[32:48].astype(np.int64)
            npch = int(h[P])
            assert 1 <= npch <= 16 * SUB
            used = {word_index(p, SUB) for p in range(npch)}
            for wi in range(32):
                assert bool(t[wi] & VALID) == (wi in used)
            # header: chunk counts per (sub, half) and the MMA N
            for sb in range(SUB):
                for hh in range(2):
                    want = sum(1 for p in range(npch) if p % SUB == sb and (p // SUB) % 2 == hh)
                    assert h[NCH + sb * 2 + hh] == want
            n0 = max(h[NCH + sb * 2] for sb in range(SUB))
            n1 = max(h[NCH + sb * 2 + 1] for sb in range(SUB))
            assert h[N] % 16 == 0 and 16 <= h[N] <= 128 and h[N] >= (64 + 8 * n1 if n1 else 8 * n0)
            run_key, run_open = None, False
            for p in range(npch):
                w = int(t[word_index(p, SUB)])
                gid, qb, cnt = w & 1023, (w >> 10) & 255, (w >> 18) & 15
                assert 1 <= cnt <= 8
                if r == 1:                                   # dense: 8 consecutive primal rows
                    assert qb == 0
                    for u in range(cnt):
                        assert (gid + u, 0) not in seen
                        seen.add((gid + u, 0))
                    continue
                assert qb >= 1 and (qb - 1) % 7 == 0 and cnt == 1 + min(7, r - qb)
                assert bool(w & OWNER) == (qb == 1)
                if w & OWNER:
                    assert gid not in owners
                    owners.add(gid)
                    seen.add((gid, 0))
                for u in range(1, cnt):
                    assert (gid, qb + u - 1) not in seen
                    seen.add((gid, qb + u - 1))
                # runs: contiguous chunks of one (receiver, slot group)
                i = gid // nb if kind >= FIRST else gid
                key = (i, qb) if kind >= FIRST else (gid, qb)
                if run_open:
                    assert key == run_key
                run_key, run_open = key, not (w & RUNEND)
            assert not run_open                              # a run never crosses a tile
            # every thread group's last chunk of a run is flagged
            start = 0
            for p in range(npch):
                if int(t[word_index(p, SUB)]) & RUNEND:
                    for q in range(start, p + 1):
                        flagged = bool(int(t[word_index(q, SUB)]) & GRPEND)
                        assert flagged == (kind == FIRST or q + 2 * SUB > p)      # first block: every chunk flushes
                    start = p + 1
            if kind in (FIRST, MID):
                rows = h[WIN_NR] * h[WIN_NS]
                assert rows <= h[MR] and 8 <= h[MR]
                assert h[WIN_I] <= h[IFIRST] <= h[ILAST] < h[WIN_I] + h[WIN_NR]
                if kind == MID:
                    for p in range(npch):
                        w = int(t[word_index(p, SUB)])
                        qb, cnt = (w >> 10) & 255, (w >> 18) & 15
                        lo = 0 if (w & OWNER) else qb
                        assert h[WIN_S] <= lo and qb + cnt - 2 < h[WIN_S] + h[WIN_NS]
        assert len(seen) == ngroups * r, (kind, len(seen), ngroups * r)
        if r > 1:
            assert len(owners) == ngroups
        if kind in (FIRST, MID, EDGE1):
            assert tab[-1][32 + FLUSH] == 1


def test_tensor_core_eligibility():
    """(128, 64) and (64, 32) run on the tcgen05 engine for every shipped system size; (256, 32) does not."""
    for n, dim, U, H, want in [(13, 3, 128, 64, True), (4, 2, 128, 64, True), (22, 3, 64, 32, True), (19, 3, 256, 32, False)]:
        eng = Engine(CnfConfig(n, dim, 0.01, 1.0, 3, (U,) * 2, H, 8, 1))
        assert (eng.lib.ecnf_solve_tensor_flops_per_eval(eng.handle) > 0) == want, (n, dim, U, H)
