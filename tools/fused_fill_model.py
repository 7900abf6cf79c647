"""CPU model of row-block fused kernels on the AM-shaped benchmark graph: how full are single-relation MMA tiles?

A fused forward kernel that keeps a block of H output rows in shared memory walks the block's edges grouped by
relation and pads every (block, relation) run to whole tiles of T entries (one relation per tensor-core tile).  The
fill = real edges / padded entries bounds its efficiency; the number of runs bounds the per-run fixed cost (weight
fragment reload).  This script prints both for several H and T on the synthetic graph bench.py uses, plus the same
for a two-level variant in which runs shorter than T/2 are routed to a scalar (FMA) tail instead of being padded.

    python tools/fused_fill_model.py [scale]
"""
import sys

import numpy as np
import torch

sys.path.insert(0, __file__.rsplit('/', 2)[0])
from torch_rgcn_b200.synthetic import SHAPES, random_triples          # noqa: E402


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    N, R, E = SHAPES['am']
    N, E = int(N * scale), int(E * scale)
    t = random_triples(N, R, E, seed=0).numpy()
    s = np.concatenate([t[:, 0], t[:, 2], np.arange(N)])
    p = np.concatenate([t[:, 1], t[:, 1] + R, np.full(N, 2 * R)])
    nnz, Rp = len(s), 2 * R + 1
    print(f'AM-shaped graph, scale {scale}: N={N} R\'={Rp} nnz={nnz} ({nnz / N:.2f} edges per row)')
    print(f'{"H":>6} {"T":>3} {"runs":>10} {"edges/run":>9} {"fill":>6} {"tiles/edge":>10} | '
          f'{"tail edges":>10} {"fill w/o tail":>13}')
    for H in (128, 256, 512, 1024, 2048, 4096):
        key = (s // H).astype(np.int64) * Rp + p
        _, cnt = np.unique(key, return_counts=True)
        for T in (8, 16, 32):
            padded = ((cnt + T - 1) // T * T).sum()
            short = cnt < T // 2                                     # runs sent to a scalar tail instead of a padded tile
            tail = cnt[short].sum()
            rest = cnt[~short]
            padded2 = ((rest + T - 1) // T * T).sum()
            print(f'{H:6d} {T:3d} {len(cnt):10d} {cnt.mean():9.2f} {nnz / padded:6.3f} {padded / T / nnz:10.4f} | '
                  f'{tail / nnz:10.3f} {rest.sum() / max(padded2, 1):13.3f}')
    # accumulator footprint: H rows x 64 fp32 columns
    for H in (256, 512, 1024, 2048):
        print(f'H={H}: accumulators {H * 64 * 4 / 1024:.0f} KB of shared memory'
              f'{" (needs a 2-CTA cluster / DSMEM)" if H * 256 > 200 * 1024 else ""}')


if __name__ == '__main__':
    main()
