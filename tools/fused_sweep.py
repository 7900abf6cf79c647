"""Time the am64 layer (forward / backward, CUDA events, L2 flushed between steps) for a list of fused-kernel
settings.  Usage: python tools/fused_sweep.py [0 640 512/s4 320/c2 640/bulk ...]
  0 = two-phase kernels; ROWS[/sSTAGES][/bulk][/f] = fused row-block kernel (f: forward only, two-phase backward)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from torch_rgcn_b200.layers import RelationalGraphConvolutionNC   # noqa: E402
from torch_rgcn_b200.synthetic import SHAPES, random_triples      # noqa: E402
from torch_rgcn_b200.utils import add_inverse_and_self             # noqa: E402


def main():
    settings = sys.argv[1:] or ['0', '512', '384', '256']
    skew = os.environ.get('SWEEP_SKEW', '0') == '1'
    dev = torch.device('cuda:0')
    N, R, E = SHAPES['am']
    t = random_triples(N, R, E, seed=0, device=dev, rel_dist='zipf' if skew else 'uniform', node_skew=skew)
    tp = add_inverse_and_self(t, N, R, device=dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    X = torch.randn(N, 64, device=dev, generator=gen).to(torch.bfloat16)
    G = torch.randn(N, 64, device=dev, generator=gen)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ref = None
    for s in settings:
        parts = s.split('/')
        os.environ['RGCN_FUSED'] = '0' if parts[0] == '0' else ('1' if 'f' in parts[1:] else '2')
        for k in ('RGCN_FUSED_STAGES', 'RGCN_FUSED_CTAS', 'RGCN_FUSED_TMA'):
            os.environ.pop(k, None)
        if parts[0] != '0':
            os.environ['RGCN_FUSE_ROWS'] = parts[0]
        for q in parts[1:]:
            if q == 'f':
                pass
            elif q == 'bulk':
                os.environ['RGCN_FUSED_TMA'] = 'bulk'
            elif q[0] == 's':
                os.environ['RGCN_FUSED_STAGES'] = q[1:]
            elif q[0] == 'c':
                os.environ['RGCN_FUSED_CTAS'] = q[1:]
        torch.manual_seed(2)
        layer = RelationalGraphConvolutionNC(triples=tp, num_nodes=N, num_relations=2 * R + 1, in_features=64,
                                             out_features=64, decomposition={'type': 'block', 'num_blocks': 4}).to(dev)
        tf = tb = 0.0
        steps, warm = 10, 3
        for k in range(steps + warm):
            flush.zero_()
            x = X.detach().requires_grad_(True)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record()
            out = layer(x)
            e[1].record()
            out.backward(G)
            e[2].record()
            torch.cuda.synchronize()
            if k >= warm:
                tf += e[0].elapsed_time(e[1]) / steps
                tb += e[1].elapsed_time(e[2]) / steps
        plan = layer._plan_cache[1]
        info = ''
        if plan.fuse_rows:
            metas = [a['meta'].tolist() for a in plan._fused if a is not None]
            info = f' fused_ok={plan.fused_ok} meta(items,tiles,overflow,split,flagged)={[m[:5] for m in metas]} fill={plan.nnz / (16.0 * metas[0][1]):.3f}'
        if ref is None:
            ref = (out.detach().clone(), x.grad.detach().float().clone())
            diff = ''
        else:
            diff = (f' max|out-ref|={(out.detach() - ref[0]).abs().max().item():.4g} (scale {ref[0].abs().max().item():.3g})'
                    f' max|gx-ref|={(x.grad.float() - ref[1]).abs().max().item():.4g} (scale {ref[1].abs().max().item():.3g})')
        print(f'setting={s} skew={skew} fwd {tf:.3f} ms bwd {tb:.3f} ms  {plan.nnz / (tf + tb) / 1e6:.2f} G edges/s{info}{diff}',
              flush=True)
        del layer, plan


if __name__ == '__main__':
    main()
