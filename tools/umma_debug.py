"""Diagnostics for the tcgen05 gathered GEMM: a single relation, structured operands, error pattern by row / column."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import numpy as np
import torch
from torch_rgcn_b200 import GraphPlan, rgcn_propagate, _lib

dev = torch.device('cuda:0')
I, O = int(sys.argv[1]) if len(sys.argv) > 1 else 64, int(sys.argv[2]) if len(sys.argv) > 2 else 64
E = int(sys.argv[3]) if len(sys.argv) > 3 else 128
N = 512
rng = np.random.RandomState(0)
tp = np.stack([rng.randint(0, N, E), np.zeros(E, np.int64), rng.randint(0, N, E)], 1)
plan = GraphPlan(torch.as_tensor(tp).to(dev), N, 1, _lib.NORM_ROW)
val = plan.val[:E].cpu().numpy().astype(np.float64)
X = torch.randn(N, I, device=dev).to(torch.bfloat16)
W = torch.randn(1, I, O, device=dev)
out = rgcn_propagate(plan, 'dense', I, O, X, weights=W)
torch.cuda.synchronize()
Xn = X.float().cpu().numpy().astype(np.float64)
Wb = W.to(torch.bfloat16).float().cpu().numpy().astype(np.float64)
ref = np.zeros((N, O))
np.add.at(ref, tp[:, 0], val[:, None] * (Xn[tp[:, 2]] @ Wb[0]))
got = out.cpu().numpy()
err = np.abs(got - ref)
print('I O E', I, O, E, 'max err', err.max(), 'scale', np.abs(ref).max(), 'nonzero rows got/ref', int((np.abs(got).sum(1) > 0).sum()), int((np.abs(ref).sum(1) > 0).sum()))
if err.max() > 1e-4 * np.abs(ref).max():
    bad_cols = np.flatnonzero(err.max(0) > 1e-4 * np.abs(ref).max())
    bad_rows = np.flatnonzero(err.max(1) > 1e-4 * np.abs(ref).max())
    print('bad cols', len(bad_cols), bad_cols[:40]); print('bad rows', len(bad_rows), bad_rows[:20])
    r = bad_rows[0]
    print('row', r, 'got', got[r, :8], 'ref', ref[r, :8])
