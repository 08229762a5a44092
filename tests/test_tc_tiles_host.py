"""Host-side invariants of the tensor-core engine's tile tables (ecnf_solve_tc.cuh: tc_pack), no GPU needed.

The epilogue threads rely on them: every (group, slot) row appears exactly once (plus flagged repeats of a primal),
segments are 'primal column, then its tangent columns' starting on 8-column chunk boundaries inside one 64-column half,
the group words invert the column map, the header masks agree with the column words, and tiles of the message-passing
kinds never span two receiver windows."""
import ctypes as C

import numpy as np
import pytest

from ecnf_b200.engine import CnfConfig, Engine

TILE_WORDS = 240
VALID, PRIMAL, DUP = 1 << 31, 1 << 30, 1 << 29
NC0, NC1, N, G0, NG, WIN, FLUSH, IFIRST, ILAST, SEGS = range(10)


def table(eng, kind):
    cnt = eng.lib.ecnf_solve_tc_tile_table(eng.handle, kind, None, 0)
    buf = np.zeros(cnt * TILE_WORDS, np.uint32)
    got = eng.lib.ecnf_solve_tc_tile_table(eng.handle, kind, buf.ctypes.data_as(C.POINTER(C.c_uint32)), buf.size)
    assert got == cnt
    return buf.reshape(cnt, TILE_WORDS)


@pytest.mark.parametrize("n,dim", [(13, 3), (4, 2), (2, 2), (7, 3)])
def test_tile_tables(n, dim):
    eng = Engine(CnfConfig(n, dim, 0.01, 1.0, 3, (128, 128, 128), 64, 8, 1))
    D, ND = n * dim, 1 + n * dim
    rs = {0: 1, 1: ND, 2: 1 + 2 * dim, 3: ND, 4: 1 + dim, 5: 1}     # 5 = primal-only edge rows (sample_cnf, no divergence)
    for kind in range(6):
        tab = table(eng, kind)
        assert len(tab) > 0
        r, ngroups = rs[kind], (n if kind < 2 else n * (n - 1))
        seen = set()
        for t in tab:
            cols, grp, h = t[:128], t[128:192], t[192:208].astype(np.int64)
            assert h[N] % 16 == 0 and 16 <= h[N] <= 128
            assert h[N] >= (64 + h[NC1] if h[NC1] else h[NC0])
            segs = 0
            for half in range(2):
                cur_g, cur_pc, last_q = -1, -1, -1
                for c in range(64):
                    w = int(cols[64 * half + c])
                    if not (w & VALID):
                        assert w == 0
                        continue
                    assert c < h[NC0 + half]
                    g, q, pc = w & 1023, (w >> 10) & 255, (w >> 18) & 63
                    if w & PRIMAL:
                        assert q == 0 and pc == c
                        if kind not in (0, 5):
                            assert c % 8 == 0                      # chunk-aligned segment start
                            segs |= 1 << (8 * half + c // 8)
                        cur_g, cur_pc, last_q = g, c, 0
                    else:
                        assert g == cur_g and pc == cur_pc and q > last_q   # tangents follow their primal, same half
                        last_q = q
                    if not (w & DUP):
                        assert (g, q) not in seen
                        seen.add((g, q))
                        lg = g - h[G0]
                        assert 0 <= lg < h[NG]
                        if kind in (0, 5):
                            assert lg == 64 * half + c     # dense packing: column = local group index
                        if lg < 64:
                            gw = int(grp[lg])
                            qs, ca, cb = gw & 255, (gw >> 8) & 255, (gw >> 16) & 255
                            assert (ca + q if q < qs else cb + 1 + q - qs) == 64 * half + c
                    mask = int(t[224 + (64 * half + c) // 32])
                    assert bool(mask >> ((64 * half + c) % 32) & 1) == bool(w & PRIMAL)
            if kind not in (0, 5):
                assert h[SEGS] == segs
            if kind in (2, 3):   # one receiver window per tile
                rw = max(1, 40 // ND)
                rw = min(rw, n)
                assert h[IFIRST] // rw == h[ILAST] // rw == h[WIN] // rw and h[WIN] % rw == 0
        assert len(seen) == ngroups * r
        if kind in (2, 3, 5):
            assert tab[-1][192 + FLUSH] == 1
