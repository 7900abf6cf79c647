"""numpy restatement of the fused row-block lists (include/rgcn_b200.h: rgcn_fused) and of what the fused kernel
does with them.  Test infrastructure: the GPU plan is compared with `expected_lists` entry by entry, and
`emulate_forward` ties the list definition back to the oracle on the CPU."""
import numpy as np

AHEAD = 8          # RGCN_FUSE_AHEAD
REC_WORDS = 36     # RGCN_FUSE_REC_WORDS


def expected_lists(tp, N, Rp, val, FR, item_tiles, backward):
    """numpy restatement of include/rgcn_b200.h: rgcn_fused for one direction."""
    s, p, o = tp[:, 0], tp[:, 1], tp[:, 2]
    a, other = (o, s) if backward else (s, o)
    blk = a // FR
    key = ((blk * Rp + p) * 2 + a % 2) * N + a
    order = np.argsort(key, kind='stable')
    grp = (key // N // 2)[order]
    starts = np.flatnonzero(np.r_[True, grp[1:] != grp[:-1]])
    ends = np.r_[starts[1:], len(order)]
    ntile = (ends - starts + 15) // 16
    tbase = np.r_[0, np.cumsum(ntile)[:-1]]
    total = int(ntile.sum())
    col = np.full(total * 16, -1, np.int64)
    row = np.zeros(total * 16, np.int64)
    v = np.zeros(total * 16, np.float32)
    tile_rel = np.zeros(total, np.int64)
    rank = np.zeros(total * 16, np.int64)
    NB = (N + FR - 1) // FR
    run_blk = grp[starts] // Rp
    flags = np.zeros(total, bool)
    for g in range(len(starts)):
        e = order[starts[g]:ends[g]]
        n = len(e)
        i = np.arange(n)
        tile_i, w = i % ntile[g], i // ntile[g]            # dealt round-robin over the run's tiles
        m = (n - tile_i + ntile[g] - 1) // ntile[g]         # entries of that tile
        h = (m + 1) // 2
        slot = np.where(w < h, 2 * w, 2 * (w - h) + 1)      # first half -> even slots, second half -> odd slots
        pos = (tbase[g] + tile_i) * 16 + slot
        for tl in range(ntile[g]):                          # tiles in which two entries share a row
            sel = np.flatnonzero(tile_i == tl)
            rows_here = a[e[sel]]
            flags[tbase[g] + tl] = len(np.unique(rows_here)) < len(rows_here)
            assert sorted(slot[sel]) == list(range(len(sel)))
            seen = {}
            for k in sel:                                   # rank of an entry among the tile's entries of its row
                rank[pos[k]] = seen.get(a[e[k]], 0)
                seen[a[e[k]]] = rank[pos[k]] + 1
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
    return dict(col=col, row=row, val=v, rank=rank, tile_rel=tile_rel, serial=flags, blk_tile=blk_tile, items=np.array(items, np.int64),
                total=total, split=split)


def records(lists):
    """The (tiles, REC_WORDS) int32 record array the plan holds for `lists` (rgcn_fused.rec)."""
    total = lists['total']
    rec = np.zeros((total, REC_WORDS), np.int32)
    slot = np.arange(16)
    word = 4 * (slot % 8) + 2 * (slot // 8)
    rows = lists['row'].reshape(total, 16)
    ranks = lists['rank'].reshape(total, 16)
    rec[:, word] = (rows * 256 + (rows & 1) * 64 + ranks).astype(np.int32)
    rec[:, word + 1] = lists['val'].reshape(total, 16).view(np.int32)
    rec[:, word] *= (rec[:, word + 1] != 0) | (lists['col'].reshape(total, 16) >= 0)     # padding stays all-zero
    ahead = np.minimum(np.arange(total) + AHEAD, total - 1)
    ahead = np.where(np.arange(total) + AHEAD < total, ahead, np.arange(total))
    maxrank = ranks.max(axis=1)
    assert ((maxrank > 0) == lists['serial']).all()
    rec[:, 32] = lists['tile_rel'][ahead].astype(np.int32) | (maxrank << 24).astype(np.int32)
    rec[:, 33] = lists['tile_rel']
    rec[:, 34] = maxrank
    return rec


def emulate_forward(lists, N, FR, X, blocks, bias):
    """out[s] = bias + sum_e val_e X[o_e] blockdiag(blocks[p_e]) computed the way k_rowblock walks the lists:
    per work item a (FR, O) tile of sums, one relation per 16-entry tile, padding skipped, split items added."""
    Rp, nb, bi, bo = blocks.shape
    O = nb * bo
    out = np.zeros((N, O))
    seen = np.zeros(N, bool)
    for b, t0, t1, shared in lists['items']:
        tile = np.zeros((FR, O))
        for ti in range(t0, t1):
            W = blocks[lists['tile_rel'][ti]]
            ent = np.arange(ti * 16, ti * 16 + 16)
            ent = ent[lists['val'][ent] != 0]
            if len(ent) == 0:
                continue
            x = X[lists['col'][ent]].reshape(len(ent), nb, bi)
            msg = lists['val'][ent, None] * np.einsum('ebi,bio->ebo', x, W).reshape(len(ent), O)
            rows = lists['row'][ent]
            for q in range(int(lists['rank'][ent].max()) + 1):   # rank by rank; within a rank a read-modify-write of all
                m = lists['rank'][ent] == q                      # entries at once: equal rows would lose updates,
                assert len(np.unique(rows[m])) == m.sum()        # which the ranks must prevent
                tile[rows[m]] = tile[rows[m]] + msg[m]
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
