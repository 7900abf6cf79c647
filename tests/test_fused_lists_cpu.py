"""CPU check of the fused row-block list definition: walking the lists the way the kernel does reproduces the
oracle (forward with the destination lists, feature gradient with the source lists)."""
import numpy as np
import pytest

from fused_ref import emulate_forward, expected_lists, records
from oracle import rgcn_oracle as orc


@pytest.mark.parametrize('FR,item_tiles', [(64, 512), (32, 3)])
def test_walking_the_lists_reproduces_the_oracle(FR, item_tiles):
    rng = np.random.RandomState(5)
    N, R, E = 300, 3, 2500
    t = np.stack([rng.randint(0, N, E) ** 2 % N, rng.randint(0, R, E), rng.randint(0, N, E) % 20], 1)   # hubs + multi-edge segments
    tp = orc.add_inverse_and_self(t, N, R)
    Rp = 2 * R + 1
    blocks = rng.randn(Rp, 4, 16, 16).astype(np.float32)
    bias = rng.randn(64).astype(np.float32)
    X = rng.randn(N, 64).astype(np.float32)
    G = rng.randn(N, 64).astype(np.float32)
    ref_out, ref_g = orc.nc_layer(tp, N, Rp, {'blocks': blocks, 'bias': bias}, X, True, G)
    val = orc.nc_edge_values(tp, N, Rp, True)
    fwd = expected_lists(tp, N, Rp, val, FR, item_tiles, backward=False)
    assert fwd['split'] > 0 or item_tiles == 512
    assert fwd['serial'].any() and not fwd['serial'].all()
    np.testing.assert_allclose(emulate_forward(fwd, N, FR, X, blocks, bias), ref_out, atol=1e-4, rtol=1e-4)
    bwd = expected_lists(tp, N, Rp, val, FR, item_tiles, backward=True)
    gx = emulate_forward(bwd, N, FR, G, blocks.transpose(0, 1, 3, 2), np.zeros(64))
    np.testing.assert_allclose(gx, ref_g['features'], atol=1e-4, rtol=1e-4)


def test_record_layout():
    """Slots pair an even with an odd row where a run has both, and padding records are all-zero."""
    rng = np.random.RandomState(2)
    N, Rp = 200, 3
    tp = np.stack([rng.randint(0, N, 900), rng.randint(0, Rp, 900), rng.randint(0, N, 900)], 1)
    val = np.ones(len(tp), np.float32)
    lists = expected_lists(tp, N, Rp, val, 64, 512, backward=False)
    rec = records(lists)
    rows, vals = lists['row'].reshape(-1, 16), lists['val'].reshape(-1, 16)
    pairs = both = 0
    for t in range(lists['total']):
        m = int((vals[t] != 0).sum())
        assert (vals[t][:m] != 0).all() and (rec[t].reshape(-1)[:32].reshape(8, 2, 2)[:, :, 1] != 0).sum() == m
        for q in range(m // 2):
            pairs += 1
            both += (rows[t, 2 * q] % 2) != (rows[t, 2 * q + 1] % 2)
    assert both / pairs > 0.8, (both, pairs)
