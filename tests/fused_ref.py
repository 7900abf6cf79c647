"""numpy restatement of the fused row-block lists (include/rgcn_b200.h: rgcn_fused) and of what the fused kernel
does with them.  Test infrastructure: the GPU plan is compared with `expected_lists` entry by entry, and
`emulate_forward` ties the list definition back to the oracle on the CPU."""
import numpy as np


def expected_lists(tp, N, Rp, val, FR, item_tiles, backward, order=1):
    """numpy restatement of include/rgcn_b200.h: rgcn_fused for one direction."""
    s, p, o = tp[:, 0], tp[:, 1], tp[:, 2]
    a, other = (o, s) if backward else (s, o)
    blk = a // FR
    key = ((blk * Rp + p) * 4 + (a % 4 if order else 0)) * N + a
    order_mode, order = order, np.argsort(key, kind='stable')
    grp = (key // N // 4)[order]
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
    flags = np.zeros(total, bool)
    for g in range(len(starts)):
        e = order[starts[g]:ends[g]]
        i = np.arange(len(e))
        if order_mode:                      # dealt round-robin over the run's tiles, alternating between the halves
            w = i // ntile[g]
            tile_i, slot = i % ntile[g], (w % 2) * 8 + (w // 2 % 2) * 4 + w // 4
        else:
            tile_i, slot = i // 16, i % 16
        pos = (tbase[g] + tile_i) * 16 + slot
        for tl in range(ntile[g]):          # tiles in which two entries added in the same step share a row
            for half in ((0, 1) if order_mode else (None,)):
                sel = (tile_i == tl) if half is None else (tile_i == tl) & (slot // 8 == half)
                rows_here = a[e[sel]]
                flags[tbase[g] + tl] |= len(np.unique(rows_here)) < len(rows_here)
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
    return dict(col=col, row=row, val=v, tile_rel=tile_rel, serial=flags, blk_tile=blk_tile, items=np.array(items, np.int64),
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
            for half in (0, 1):                          # the kernel adds slots 0-7, then 8-15, each as one step
                ent = np.arange(ti * 16 + 8 * half, ti * 16 + 8 * half + 8)
                ent = ent[lists['val'][ent] != 0]
                if len(ent) == 0:
                    continue
                x = X[lists['col'][ent]].reshape(len(ent), nb, bi)
                msg = lists['val'][ent, None] * np.einsum('ebi,bio->ebo', x, W).reshape(len(ent), O)
                rows = lists['row'][ent]
                if lists['serial'][ti]:
                    np.add.at(tile, rows, msg)
                else:                                    # read-modify-write of all entries at once: equal rows would
                    tile[rows] = tile[rows] + msg        # lose updates, which is what the plan's flag must prevent
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
