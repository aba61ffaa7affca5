"""The stream-K decomposition of csrc/gemm_dmma.cuh (dgemm_tn_streamk_kernel::decode_run, streamk_reduce_kernel) restated in
Python: for any tile count, k-tile count and grid, the CTAs' runs cover every (tile, k-tile) exactly once, a tile's pieces
sit in consecutive workspace slots 0 .. last - first in CTA order (what the fixed-order reduction relies on), complete tiles
are exactly the ones with a single piece, and the slot bound the host planner allocates for holds."""
import itertools

import pytest


def runs(tiles, KT, G):
    """Pieces (cta, tile, kt_begin, nkt, slot) as the kernel derives them; P = k-tiles per CTA as the host sets it."""
    U = tiles * KT
    P = -(-U // G)
    grid = -(-U // P)
    out = []
    for cta in range(grid):
        lo, hi = cta * P, min((cta + 1) * P, U)
        w = 0
        while True:
            tile = lo // KT + w
            tb = tile * KT
            kt_begin = lo - tb if w == 0 else 0
            nkt = min(KT, hi - tb) - kt_begin
            if nkt <= 0 or tile >= tiles:
                break
            out.append((cta, tile, kt_begin, nkt, cta - tb // P))
            w += 1
    return out, P


@pytest.mark.parametrize("tiles,KT,G", [(157, 313, 148), (40, 1250, 148), (1, 1250, 148), (20, 313, 148), (80, 49, 148),
                                         (6280, 16, 148), (3, 7, 148), (149, 5, 148)] +
                         [(t, k, g) for t, k, g in itertools.product((1, 2, 37, 148, 300), (4, 48, 1000), (1, 7, 148))])
def test_runs_tile_the_work_exactly(tiles, KT, G):
    if tiles * KT // 4 < 1:
        pytest.skip("fewer than four k-tiles in total: the planner never picks stream-K")
    G = min(G, max(1, tiles * KT // 4))
    pieces, P = runs(tiles, KT, G)
    covered = {}
    for cta, tile, b, n, slot in pieces:
        assert 0 <= tile < tiles and n > 0 and 0 <= b and b + n <= KT
        for kt in range(b, b + n):
            assert (tile, kt) not in covered, "k-tile done twice"
            covered[(tile, kt)] = cta
    assert len(covered) == tiles * KT
    slots_bound = -(-KT // P) + 1                              # what gemm_host.inl reserves per element
    by_tile = {}
    for cta, tile, b, n, slot in pieces:
        by_tile.setdefault(tile, []).append((cta, b, n, slot))
    for tile, ps in by_tile.items():
        first, last = tile * KT // P, ((tile + 1) * KT - 1) // P       # the reduce kernel's view
        assert [p[0] for p in ps] == list(range(first, last + 1)), "pieces of a tile come from consecutive CTAs"
        assert [p[3] for p in ps] == list(range(0, last - first + 1)), "slots are 0 .. last - first in CTA order"
        assert len(ps) <= slots_bound
        complete = [p for p in ps if p[2] == KT]
        assert (len(ps) == 1) == (len(complete) == 1), "a tile is stored directly iff it has a single piece"
        assert sorted(p[1] for p in ps) == [p[1] for p in ps]          # ascending k: the summation order is the k order
    # balance: no CTA has more than P k-tiles, all but the last have exactly P
    per_cta = {}
    for cta, tile, b, n, slot in pieces:
        per_cta[cta] = per_cta.get(cta, 0) + n
    last_cta = max(per_cta)
    assert all(v == P for c, v in per_cta.items() if c != last_cta) and 0 < per_cta[last_cta] <= P
