"""Multi-GPU parity of the relation-sharded layer (run under torchrun on N >= 2 GPUs, NCCL):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/sharded_check.py

Every rank builds the same layer, runs the single-GPU engine on the full graph and the sharded engine on its relations,
and checks that output, feature gradient and (after sync_parameter_grads) parameter gradients agree.
SHARD=rows in the environment checks the experimental row-sharded layer (parallel.RowShardedNC) instead.
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from torch_rgcn_b200.layers import RelationalGraphConvolutionNC
    from torch_rgcn_b200.parallel import RelationShardedNC, RowShardedNC
    Sharded = RowShardedNC if os.environ.get('SHARD') == 'rows' else RelationShardedNC
    from torch_rgcn_b200.synthetic import random_triples
    from torch_rgcn_b200.utils import add_inverse_and_self
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()
    N, R, E = 50000, 23, 400000
    cases = [
        dict(in_f=16, out_f=16, decomp={'type': 'block', 'num_blocks': 2}, vertical=False, dtype=torch.float32),
        dict(in_f=16, out_f=8, decomp={'type': 'basis', 'num_bases': 5}, vertical=True, dtype=torch.float32),
        dict(in_f=64, out_f=64, decomp={'type': 'block', 'num_blocks': 4}, vertical=False, dtype=torch.bfloat16),
        dict(in_f=None, out_f=16, decomp=None, vertical=False, dtype=torch.float32),
        # generic row-sharded path of bf16 layers: fused kernels per 64-column group / tcgen05 GEMMs on the local plans,
        # rows exchanged through symmetric memory (different widths per direction in the second case)
        dict(in_f=128, out_f=128, decomp={'type': 'block', 'num_blocks': 8}, vertical=True, dtype=torch.bfloat16),
        dict(in_f=64, out_f=128, decomp=None, vertical=False, dtype=torch.bfloat16),
    ]
    for ci, c in enumerate(cases):
        t = random_triples(N, R, E, seed=ci, device=dev, rel_dist='zipf')
        tp = add_inverse_and_self(t, N, R, device=dev)
        torch.manual_seed(100 + ci)
        ref = RelationalGraphConvolutionNC(triples=tp, num_nodes=N, num_relations=2 * R + 1, in_features=c['in_f'],
                                           out_features=c['out_f'], decomposition=c['decomp'],
                                           vertical_stacking=c['vertical']).to(dev)
        torch.manual_seed(100 + ci)
        lay = RelationalGraphConvolutionNC(triples=tp, num_nodes=N, num_relations=2 * R + 1, in_features=c['in_f'],
                                           out_features=c['out_f'], decomposition=c['decomp'],
                                           vertical_stacking=c['vertical']).to(dev)
        with torch.no_grad():
            ref.bias.normal_()
            lay.bias.copy_(ref.bias)
        sh = Sharded(lay)
        g = torch.Generator(device=dev).manual_seed(7)
        x1 = x2 = None
        if c['in_f'] is not None:
            x = torch.randn(N, c['in_f'], device=dev, generator=g).to(c['dtype'])
            x1, x2 = x.clone().requires_grad_(True), x.clone().requires_grad_(True)
        o1 = ref(x1) if x1 is not None else ref()
        G = torch.randn(o1.shape, device=dev, generator=g)
        for _ in range(3):                                   # several steps: the exchange buffers alternate
            for prm in lay.parameters():
                prm.grad = None
            if x2 is not None:
                x2.grad = None
            o2 = sh(x2) if x2 is not None else sh()
            o2.backward(G)
        o1.backward(G)
        sh.sync_parameter_grads()
        tol = dict(atol=2e-4, rtol=2e-4) if c['dtype'] == torch.float32 else dict(atol=3e-2, rtol=3e-2)
        torch.testing.assert_close(o2, o1, **tol)
        if x1 is not None:
            torch.testing.assert_close(x2.grad.float(), x1.grad.float(), **tol)
        for (n1, p1), (n2, p2) in zip(ref.named_parameters(), lay.named_parameters()):
            scale = p1.grad.abs().max().item() + 1e-12
            err = (p1.grad - p2.grad).abs().max().item() / scale
            assert err < (2e-3 if c['dtype'] == torch.float32 else 3e-2), (ci, n1, err)
        if rank == 0:
            print(f'case {ci} ok on {world} ranks', flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
