"""GPU parity of the DistMult decoder kernels and the negative-sampling corruption step (through the C ABI)."""
import numpy as np
import pytest
import torch

from conftest import load_golden, golden_names
from oracle import distmult_oracle as dm

pytestmark = pytest.mark.gpu
ATOL = 1e-4


def _dec(meta, params, dev):
    from torch_rgcn_b200.layers import DistMult
    dec = DistMult(meta['R'], meta['d'], meta['N'], meta['R'], b_init=meta['b_init']).to(dev)
    with torch.no_grad():
        for n, p in dec.named_parameters():
            p.copy_(torch.as_tensor(params[n]))
    return dec


@pytest.mark.parametrize('name', golden_names('distmult_'))
def test_distmult_matches_reference_fixture(cuda_device, name):
    meta, d, params, grads = load_golden(name)
    dec = _dec(meta, params, cuda_device)
    nodes = torch.as_tensor(d['nodes']).to(cuda_device).requires_grad_(True)
    triples = torch.as_tensor(d['triples']).to(cuda_device)
    out = dec(triples, nodes)
    assert out.shape == d['out'].shape and out.dtype == torch.float32
    np.testing.assert_allclose(out.detach().cpu().numpy(), d['out'], atol=ATOL, rtol=1e-4)
    out.backward(torch.as_tensor(d['G']).to(cuda_device))
    np.testing.assert_allclose(nodes.grad.cpu().numpy(), grads['nodes'], atol=ATOL, rtol=1e-4)
    for n, p in dec.named_parameters():
        np.testing.assert_allclose(p.grad.cpu().numpy(), grads[n], atol=ATOL, rtol=1e-4, err_msg=n)
    nodes.grad = None
    dec.zero_grad()
    pen = dec.s_penalty(triples, nodes)
    assert pen.dim() == 0
    np.testing.assert_allclose(pen.item(), d['penalty'], rtol=1e-5)
    (pen * meta['penalty_grad']).backward()
    np.testing.assert_allclose(nodes.grad.cpu().numpy(), d['pgrad_nodes'], atol=1e-6, rtol=1e-4)
    np.testing.assert_allclose(dec.relations.grad.cpu().numpy(), d['pgrad_relations'], atol=1e-6, rtol=1e-4)


@pytest.mark.parametrize('N,R,dim,B,bias', [(4000, 18, 128, 60000, False), (900, 7, 50, 20000, True),
                                            (300, 3, 200, 5000, True), (64, 2, 4, 1000, False)])
def test_distmult_matches_oracle_at_size(cuda_device, N, R, dim, B, bias):
    """WN18-like width, odd widths (scalar path), heavy repetition of nodes and relations (gradient accumulation),
    and the batch layout of training (positives followed by blocks of negatives that share the relation)."""
    from torch_rgcn_b200.layers import DistMult
    g = torch.Generator().manual_seed(5)
    pos = torch.stack([torch.randint(0, N, (B // 5,), generator=g), torch.randint(0, R, (B // 5,), generator=g),
                       torch.randint(0, N, (B // 5,), generator=g)], 1)
    neg = pos[:, None, :].expand(B // 5, 4, 3).contiguous()
    neg[:, ::2, 0] = torch.randint(0, N, (B // 5, 2), generator=g)
    neg[:, 1::2, 2] = torch.randint(0, N, (B // 5, 2), generator=g)
    triples = torch.cat([pos, neg.view(-1, 3)], 0)
    torch.manual_seed(6)
    dec = DistMult(R, dim, N, R, b_init='normal' if bias else None).to(cuda_device)
    nodes = torch.randn(N, dim, generator=g).to(cuda_device).requires_grad_(True)
    G = torch.randn(triples.size(0), generator=g)
    out = dec(triples.to(cuda_device), nodes)
    out.backward(G.to(cuda_device))
    P = {n: p.detach().cpu().numpy() for n, p in dec.named_parameters()}
    biases = [P.get(k) for k in ('sbias', 'pbias', 'obias')]
    ref = dm.score(triples.numpy(), nodes.detach().cpu().numpy(), P['relations'], *biases)
    np.testing.assert_allclose(out.detach().cpu().numpy(), ref, atol=ATOL, rtol=1e-4)
    rg = dm.score_backward(triples.numpy(), nodes.detach().cpu().numpy(), P['relations'], G.numpy(), with_bias=bias)

    def close(got, want, name):                    # sums over up to B / R terms: tolerance relative to the tensor's scale
        np.testing.assert_allclose(got, want, atol=ATOL * max(1.0, np.abs(want).max()), rtol=1e-4, err_msg=name)
    close(nodes.grad.cpu().numpy(), rg['nodes'], 'nodes')
    for n, p in dec.named_parameters():
        close(p.grad.cpu().numpy(), rg[n], n)
    pen = dec.s_penalty(triples.to(cuda_device), nodes)
    np.testing.assert_allclose(pen.item(), dm.penalty(triples.numpy(), nodes.detach().cpu().numpy(), P['relations']), rtol=1e-5)


def test_distmult_edge_cases(cuda_device):
    from torch_rgcn_b200.layers import DistMult
    dec = DistMult(3, 8, 10, 3).to(cuda_device)
    nodes = torch.randn(10, 8, device=cuda_device, requires_grad=True)
    empty = dec(torch.zeros(0, 3, dtype=torch.long, device=cuda_device), nodes)
    assert empty.shape == (0,)
    empty.sum().backward()
    assert torch.all(nodes.grad == 0) and torch.all(dec.relations.grad == 0)
    with pytest.raises(IndexError):
        dec(torch.tensor([[0, 0, 10]], device=cuda_device), nodes)
    with pytest.raises(IndexError):
        dec(torch.tensor([[0, 3, 1]], device=cuda_device), nodes)
    with pytest.raises(AssertionError):
        dec(torch.zeros(2, 3, dtype=torch.int32, device=cuda_device), nodes)
    # frozen decoder: only the node gradient is produced
    dec.relations.requires_grad_(False)
    nodes.grad = None
    dec(torch.tensor([[1, 2, 3], [1, 0, 1]], device=cuda_device), nodes).sum().backward()
    assert dec.relations.grad is None or torch.all(dec.relations.grad == 0)
    assert nodes.grad.abs().sum() > 0


@pytest.mark.parametrize('name', golden_names('negsample_'))
def test_corruption_kernel_matches_reference_fixture(cuda_device, name):
    from torch_rgcn_b200 import _lib
    meta, d, _, _ = load_golden(name)
    batch = torch.as_tensor(d['batch']).to(cuda_device).contiguous()
    head = torch.as_tensor(d['head']).to(cuda_device).to(torch.uint8)
    cor = torch.as_tensor(d['corruptions']).to(cuda_device)
    _lib.check(_lib.lib.rgcn_corrupt_triples(_lib.ptr(batch), _lib.ptr(head), _lib.ptr(cor), cor.numel(), _lib.stream_ptr()))
    np.testing.assert_array_equal(batch.view(-1, 3).cpu().numpy(), d['out'])


def test_negative_sampling_draws_like_the_reference(cuda_device):
    """Same two draws in the same order as utils/misc.py:178-183 (randint, then bernoulli) on the given device, then
    the masked assignment: re-drawing them from the same seed and applying the oracle reproduces the batch."""
    from torch_rgcn_b200.decoder import negative_sampling
    bs, ns, N = 500, 10, 40943
    g = torch.Generator().manual_seed(3)
    pos = torch.stack([torch.randint(0, N, (bs,), generator=g), torch.randint(0, 18, (bs,), generator=g),
                       torch.randint(0, N, (bs,), generator=g)], 1).to(cuda_device)
    batch = pos.clone()[:, None, :].expand(bs, ns, 3).contiguous()
    before = batch.cpu().numpy().copy()
    torch.manual_seed(77)
    out = negative_sampling(batch, N, 0.5, device=cuda_device)
    assert out.shape == (bs * ns, 3) and out.data_ptr() == batch.data_ptr()      # in place, like the reference
    torch.manual_seed(77)
    cor = torch.randint(size=(bs * ns,), low=0, high=N, dtype=torch.long, device=cuda_device)
    head = torch.bernoulli(torch.empty(size=(bs, ns, 1), dtype=torch.float, device=cuda_device).fill_(0.5)).to(torch.bool)
    np.testing.assert_array_equal(out.cpu().numpy(), dm.corrupt(before, head.cpu().numpy(), cor.cpu().numpy()))
    frac_head = (out[:, 0].cpu().numpy() != before.reshape(-1, 3)[:, 0]).mean()
    assert 0.4 < frac_head < 0.6
