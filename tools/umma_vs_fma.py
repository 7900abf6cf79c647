"""Forward / feature-gradient time of a dense 512 x 512 bf16 layer: tcgen05 gathered GEMM vs the fp32 register-tiled kernel."""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
import torch
from torch_rgcn_b200 import GraphPlan, rgcn_propagate, _lib
from torch_rgcn_b200.synthetic import random_triples

dev = torch.device('cuda:0')
N, Rp, E, I, O = 250000, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000, 512, 512
tp = random_triples(N, Rp, E, seed=5, device=dev)
plan = GraphPlan(tp, N, Rp, _lib.NORM_ROW)
W = (torch.randn(Rp, I, O, device=dev) * 0.05).requires_grad_(True)
x = torch.randn(N, I, device=dev).to(torch.bfloat16).requires_grad_(True)
G = torch.randn(N, O, device=dev)
res = {}
for mode in ('1', '1m1', '0'):
    os.environ['RGCN_UMMA'] = mode[0]
    os.environ['RGCN_UMMA_WGRAD_MT'] = '1' if mode == '1m1' else '2'
    for _ in range(2):
        out = rgcn_propagate(plan, 'dense', I, O, x, weights=W)
        out.backward(G)
    torch.cuda.synchronize()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    reps = 3 if mode[0] == '1' else 1
    e0.record()
    for _ in range(reps):
        out = rgcn_propagate(plan, 'dense', I, O, x, weights=W)
    e1.record()
    for _ in range(reps):
        x.grad = None; W.grad = None
        out.backward(G, retain_graph=True)
    e2.record()
    torch.cuda.synchronize()
    res[mode] = (e0.elapsed_time(e1) / reps, e1.elapsed_time(e2) / reps, out.detach().float().clone(), W.grad.clone())
f1, b1, o1, g1 = res['1']; f0, b0, o0, g0 = res['0']
flops = 2.0 * E * I * O
print(json.dumps({'wgrad_tile_variants_fwd_bwd_ms': {k: (round(v[0], 2), round(v[1], 2)) for k, v in res.items()}}))
print(json.dumps({'edges': E, 'ms_fwd_umma': f1, 'ms_fwd_fma': f0, 'speedup_fwd': f0 / f1, 'tflops_fwd_umma': flops / f1 / 1e9,
                  'ms_bwd_umma': b1, 'ms_bwd_fma': b0,
                  'max_abs_diff': float((o1 - o0).abs().max()), 'gW_max_abs_diff': float((g1 - g0).abs().max()), 'gW_scale': float(g0.abs().max()), 'scale': float(o0.abs().max())}))
