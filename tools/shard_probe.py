"""torchrun script: kernel-level timeline of the row-sharded am64 layer step on rank 0 (torch.profiler)."""
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    from torch_rgcn_b200.layers import RelationalGraphConvolutionNC
    from torch_rgcn_b200.parallel import RowShardedNC
    from torch_rgcn_b200.synthetic import SHAPES, random_triples
    from torch_rgcn_b200.utils import add_inverse_and_self
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    dist.init_process_group('nccl', device_id=dev)
    rank = dist.get_rank()
    if os.environ.get('PROBE', 'am') == 'syn':               # quarter-scale synthetic 512-wide layer (arbitrary triples)
        N, Rp, W, nb = 1250000, 256, 512, 32
        tp = random_triples(N, Rp, 50_000_000, seed=0, device=dev)
        kw = dict(vertical_stacking=True)
    else:
        N, R, E = SHAPES['am']
        t = random_triples(N, R, E, seed=0, device=dev)
        tp = add_inverse_and_self(t, N, R, device=dev)
        Rp, W, nb, kw = 2 * R + 1, 64, 4, {}
    torch.manual_seed(2)
    layer = RowShardedNC(RelationalGraphConvolutionNC(triples=tp, num_nodes=N, num_relations=Rp, in_features=W,
                                                      out_features=W, decomposition={'type': 'block', 'num_blocks': nb},
                                                      **kw).to(dev))
    gen = torch.Generator(device=dev).manual_seed(1)
    X = torch.randn(N, W, device=dev, generator=gen).to(torch.bfloat16)
    G = torch.randn(N, W, device=dev, generator=gen)

    def step():
        x = X.detach().requires_grad_(True)
        out = layer(x)
        out.backward(G)

    reps = 3 if W > 64 else 20
    for _ in range(2 if W > 64 else 5):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    if rank == 0:
        print(f'{reps} steps: enqueue {1e3 * (t1 - t0) / reps:.3f} ms/step on the host, {1e3 * (t2 - t0) / reps:.3f} ms/step to completion')
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for _ in range(2 if W > 64 else 10):
            step()
        torch.cuda.synchronize()
    if rank == 0:
        print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=25, max_name_column_width=60))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
