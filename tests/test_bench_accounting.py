"""CPU: bench.py's roofline accounting follows SURVEY §8(d) and its workload table is well-formed."""
import importlib.util
import os

from conftest import ROOT


def _bench():
    spec = importlib.util.spec_from_file_location('bench', os.path.join(ROOT, 'bench.py'))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_algorithmic_bytes_follow_the_survey_formulas():
    b = _bench()
    from torch_rgcn_b200.synthetic import SHAPES
    N, R, E = SHAPES['am']
    Rp, nnz = 2 * R + 1, 2 * E + N
    assert (N, Rp, nnz) == (1666764, 267, 13643406)
    bf, bb, per_edge = b.algorithmic_bytes(b.WORKLOADS['am64'], N, Rp, nnz)
    # forward: one bf16 source row (128 B) + 4 B source index + 4 B segment id + 4 B explicit val per edge,
    # one fp32 output write, one weight read
    assert per_edge == 64 * 2 + 12 == 140
    w = Rp * 64 * 64 // 4
    assert bf == nnz * 140 + 4 * (N + 1) + N * 64 * 4 + 4 * w
    assert abs(bf / 1e9 - 2.3445) < 1e-3                        # the 2.34 GB of DESIGN.md §3
    assert bb == nnz * (64 * 4 + 12) + nnz * (64 * 2 + 8) + N * 64 * 4 + N * 64 * 4 + 8 * w
    assert abs(bb / 1e9 - 6.37) < 0.01
    # fp32 16 -> 16: 76 B per edge (SURVEY: 80 minus nothing implied; explicit val included)
    assert b.algorithmic_bytes(b.WORKLOADS['am16'], N, Rp, nnz)[2] == 16 * 4 + 12
    # synthetic 512-wide bf16: 1,036 B per edge
    assert b.algorithmic_bytes(b.WORKLOADS['syn'], 5000000, 256, 200000000)[2] == 1036


def test_workload_table_is_well_formed():
    b = _bench()
    from torch_rgcn_b200.synthetic import SHAPES
    kinds = {'nc', 'lp', 'lp_step', 'nc_model', 'decoder', 'ranking', 'sampling'}
    for name, wl in b.WORKLOADS.items():
        assert wl['kind'] in kinds and wl['shape'] in SHAPES and wl['label'], name
        assert wl['dtype'] in ('bf16', 'f32', 'i32'), name
    # the default line is the configuration BASELINE.json quotes its metric on (AM-shaped, block-diagonal, bf16)
    assert b.WORKLOADS['am64']['decomp'] == {'type': 'block', 'num_blocks': 4} and b.WORKLOADS['am64']['dtype'] == 'bf16'
