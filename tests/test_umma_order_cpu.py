"""Host-side restatement of the chunk decode of the tensor-core GEMM kernels (csrc/propagate_umma.cuh, umma_chunk): in
quantile order every chunk of every relation must be claimed by exactly one slot, for any relation sizes."""
import numpy as np
import pytest

CHUNK = 1024


def _decode(relptr, chunkptr, num_rels, maxc, group, slot):
    """umma_chunk(): (relation, first edge, end edge) of the slot, or None."""
    if maxc > 0:
        per = maxc * group
        g, r = divmod(slot, per)
        qq, j = divmod(r, group)
        p = g * group + j
        if p >= num_rels:
            return None
        first, n_p = chunkptr[p], chunkptr[p + 1] - chunkptr[p]
        if n_p == 0:
            return None
        q = (qq * n_p + maxc - 1) // maxc
        if q >= n_p or q * maxc // n_p != qq:
            return None
        c = first + q
    else:
        c = slot
        if c >= chunkptr[num_rels]:
            return None
        p = int(np.searchsorted(chunkptr, c, side='right') - 1)
        while chunkptr[p + 1] == chunkptr[p]:      # the device binary search lands on the last relation starting at c
            p += 1
    e0 = relptr[p] + (c - chunkptr[p]) * CHUNK
    return p, e0, min(relptr[p + 1], e0 + CHUNK)


@pytest.mark.parametrize('seed,num_rels,group', [(0, 7, 3), (1, 64, 64), (2, 33, 8), (3, 5, 1), (4, 40, 16)])
def test_every_chunk_is_claimed_exactly_once(seed, num_rels, group):
    rng = np.random.RandomState(seed)
    counts = (rng.zipf(1.3, num_rels) * 37 % 9000).astype(np.int64)
    counts[rng.randint(num_rels)] = 0                                   # an empty relation
    counts[rng.randint(num_rels)] = 20000                               # a long one
    relptr = np.concatenate([[0], np.cumsum(counts)])
    chunks = (counts + CHUNK - 1) // CHUNK
    chunkptr = np.concatenate([[0], np.cumsum(chunks)])
    maxc = int(chunks.max())
    ngroups = (num_rels + group - 1) // group
    for mc, slots in ((0, int(chunkptr[-1]) + 5), (maxc, maxc * ngroups * group)):
        seen = {}
        for slot in range(slots):
            got = _decode(relptr, chunkptr, num_rels, mc, group, slot)
            if got is None:
                continue
            p, e0, e1 = got
            assert relptr[p] <= e0 < e1 <= relptr[p + 1] and (e0 - relptr[p]) % CHUNK == 0
            assert (p, e0) not in seen, 'chunk claimed twice'
            seen[(p, e0)] = e1
        assert len(seen) == int(chunkptr[-1])
        assert sum(e1 - e0 for (p, e0), e1 in seen.items()) == int(counts.sum())
