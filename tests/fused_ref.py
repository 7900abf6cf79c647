"""numpy restatement of the fused row-block lists (include/rgcn_b200.h: rgcn_fused) and of what the fused kernel
does with them.  Test infrastructure: the GPU plan is compared with `expected_lists` entry by entry, and
`emulate_forward` ties the list definition back to the oracle on the CPU."""
import numpy as np


def expected_lists(tp, N, Rp, val, FR, item_tiles, backward):
    """numpy restatement of include/rgcn_b200.h: rgcn_fused for one direction."""
    s, p, o = tp[:, 0], tp[:, 1], tp[:, 2]
    a, other = (o, s) if backward else (s, o)
    blk = a // FR
    key = (blk * Rp + p) * N + a
    order = np.argsort(key, kind='stable')
    grp = (key // N)[order]
    starts = np.flatnonzero(np.r_[True, grp[1:] != grp[:-1]])
    ends = np.r_[starts[1:], len(order)]
    ntile = (ends - starts + 15) // 16
    tbase = np.r_[0, np.cumsum(ntile)[:-1]]
    total = int(ntile.sum())
    col = np.full(total * 16, -1, np.int64)
    row = np.zeros(total * 16, np.int64)
    v = np.zeros(total * 16, np.float32)
    tile_rel = np.zeros(total, np.int64)
    NB = (N + FR - 1) // FR
    run_blk = grp[starts] // Rp
    for g in range(len(starts)):
        e = order[starts[g]:ends[g]]
        pos = tbase[g] * 16 + np.arange(len(e))
        col[pos] = other[e]
        row[pos] = a[e] - run_blk[g] * FR
        v[pos] = val[e]
        tile_rel[tbase[g]:tbase[g] + ntile[g]] = grp[starts[g]] % Rp
    blk_tile = np.zeros(NB + 1, np.int64)
    tiles_per_blk = np.bincount(run_blk, weights=ntile, minlength=NB).astype(np.int64)
    blk_tile[1:] = np.cumsum(tiles_per_blk)
    items = []
    for b in range(NB):
        t0, t1 = blk_tile[b], blk_tile[b + 1]
        n = max(1, -(-(t1 - t0) // item_tiles))
        for i in range(n):
            items.append((b, t0 + i * item_tiles, min(t1, t0 + (i + 1) * item_tiles), int(n > 1)))
    split = int(sum(1 for b in range(NB) if blk_tile[b + 1] - blk_tile[b] > item_tiles))
    return dict(col=col, row=row, val=v, tile_rel=tile_rel, blk_tile=blk_tile, items=np.array(items, np.int64),
                total=total, split=split)


def emulate_forward(lists, N, FR, X, blocks, bias):
    """out[s] = bias + sum_e val_e X[o_e] blockdiag(blocks[p_e]) computed the way k_fused_rows walks the lists:
    per work item a zeroed (FR, O) tile, one relation per 16-entry tile, padding skipped, split items added."""
    Rp, nb, bi, bo = blocks.shape
    O = nb * bo
    out = np.zeros((N, O))
    seen = np.zeros(N, bool)
    for b, t0, t1, shared in lists['items']:
        tile = np.zeros((FR, O))
        for ti in range(t0, t1):
            W = blocks[lists['tile_rel'][ti]]
            for e in range(ti * 16, ti * 16 + 16):
                if lists['val'][e] == 0:
                    continue
                x = X[lists['col'][e]].reshape(nb, bi)
                tile[lists['row'][e]] += lists['val'][e] * np.einsum('bi,bio->bo', x, W).reshape(O)
        rows = slice(b * FR, min(N, (b + 1) * FR))
        n = rows.stop - rows.start
        if shared:
            if not seen[rows.start]:
                out[rows] = bias
            out[rows] += tile[:n]
        else:
            out[rows] = tile[:n] + bias
        seen[rows] = True
    assert seen.all()
    return out
