#!/usr/bin/env python
"""bench.py — RGCN-layer edges/sec (forward+backward) on synthetic graphs of BASELINE.json's shapes.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload am64|am16|aifb|mutag|wn18|syn] [--impl reference]

A "step" is ONE layer call: forward + backward of the RGCN layer over every edge of `triples_plus`
(metric = nnz / (t_fwd + t_bwd)).  The default workload is the AM-shaped block-diagonal
bf16 layer (BASELINE.json configs[2]) — the shape the north-star target is quoted on; the graph plan of a
node-classification layer is built once and is outside the timed region, as it is a construction-time cost
in this engine (the reference rebuilds it every forward; its port below pays that inside its timed region).

Our arm:   value = device-resident throughput (CUDA events, max over ranks); e2e = the same step through the
           layer with HOST (pinned) buffers, H2D of features + upstream gradient and D2H of output + gradients
           inside the timed region; roofline = forward gather kernel vs the measured HBM peak.
Reference: `--impl reference` times oracle/torch_sparse_port.py (a CPU port of the reference's torch.sparse
           algorithm; the pure-Python reference checkout does not travel to the GPU box) on a bounded,
           uniformly scaled-down sample of the same workload, all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: shape key, layer kind, in, out, decomposition, feature dtype, vertical, config label
    'am64': dict(shape='am', kind='nc', in_f=64, out_f=64, decomp={'type': 'block', 'num_blocks': 4}, dtype='bf16',
                 vertical=False, label='AM-shaped nc rgcn layer (1,666,764 nodes, 133 rels -> R\'=267, nnz=13,643,406), '
                                       'block-diagonal nb=4, 64->64, bf16 features / fp32 accumulate'),
    'am16': dict(shape='am', kind='nc', in_f=16, out_f=16, decomp={'type': 'block', 'num_blocks': 2}, dtype='f32',
                 vertical=False, label='AM-shaped nc rgcn layer, block-diagonal nb=2, 16->16, fp32'),
    'aifb': dict(shape='aifb', kind='nc', in_f=16, out_f=4, decomp=None, dtype='f32', vertical=True,
                 label='AIFB-shaped nc rgcn layer 2 (8,285 nodes, R\'=91, nnz=66,371), no decomposition, 16->4, fp32'),
    'mutag': dict(shape='mutag', kind='nc', in_f=16, out_f=2, decomp={'type': 'basis', 'num_bases': 30}, dtype='f32',
                  vertical=True, label='MUTAG-shaped nc rgcn layer 2 (23,644 nodes, R\'=47), basis B=30, 16->2, fp32'),
    'wn18': dict(shape='wn18', kind='lp', in_f=16, out_f=16, decomp=None, dtype='f32', vertical=False,
                 label='WN18-shaped lp rgcn layer (40,943 nodes, R\'=37, 70,721 sampled triples -> nnz=253,106), '
                       '16->16, fp32; graph build inside the step'),
    # the LP layer of configs/rgcn/lp-FB-toy.yaml (500 wide, 100 blocks of 5 x 5, dense 500 x 500 self-loop weight) on an
    # FB15k-237-shaped graph: block relations in the generic kernel, the self-loop relation as a gathered GEMM
    'fb15k_block': dict(shape='fb15k237', kind='lp', in_f=500, out_f=500, decomp={'type': 'block', 'num_blocks': 100},
                        dtype='f32', vertical=False,
                        label='FB15k-237-shaped lp rgcn layer (14,541 nodes, R\'=475, 136,057 sampled triples), block-diagonal '
                              'nb=100 (5x5) + dense self-loop weight, 500->500, fp32; graph build inside the step'),
    # the whole two-layer NodeClassifier step (featureless layer 1 -> ReLU -> layer 2 -> CE loss)
    'aifb_model': dict(shape='aifb', kind='nc_model', in_f=None, out_f=4, hidden=16, decomp=None, dtype='f32',
                       vertical=False, label='AIFB-shaped 2-layer NodeClassifier step (8,285 nodes, R\'=91, '
                                             'nnz=66,371 per layer), no decomposition, hidden 16, 4 classes, fp32'),
    'mutag_model': dict(shape='mutag', kind='nc_model', in_f=None, out_f=2, hidden=16,
                        decomp={'type': 'basis', 'num_bases': 30}, dtype='f32', vertical=False,
                        label='MUTAG-shaped 2-layer NodeClassifier step (23,644 nodes, R\'=47, nnz=172,098 per layer), '
                              'basis B=30, hidden 16, 2 classes, fp32'),
    # the DistMult decoder step of the WN18 c-rgcn config (BASELINE configs[3])
    'wn18_decoder': dict(shape='wn18', kind='decoder', in_f=128, out_f=128, decomp=None, dtype='f32', vertical=False,
                         label='WN18-shaped DistMult decoder step (40,943 nodes x 128, 18 relations; 141,442 positives + '
                               '10 negatives each = 1,555,862 scored triples): scores + L2 penalty + BCE loss, fp32'),
    # filtered ranking evaluation of the WN18 c-rgcn model (5,000 test triples, both sides)
    'wn18_ranking': dict(shape='wn18', kind='ranking', in_f=128, out_f=128, decomp=None, dtype='f32', vertical=False,
                         label='WN18-shaped filtered ranking evaluation (40,943 nodes x 128, 18 relations; 5,000 test '
                               'triples x 2 sides x 40,943 candidates; filter over 151,442 known triples), fp32'),
    # one epoch's inputs of the WN18 rgcn config (configs/rgcn/lp-WN18.yaml: 30,000 positives by
    # edge-neighbourhood sampling, 10 negatives each, general edge dropout 0.5)
    'wn18_sampling': dict(shape='wn18', kind='sampling', in_f=0, out_f=0, decomp=None, dtype='i32', vertical=False,
                          label='WN18-shaped per-step graph construction (141,442 training triples, 40,943 nodes): '
                                '30,000 edge-neighbourhood picks + 300,000 negatives + edge dropout 0.5'),
    # the whole link-prediction training step of the shipped WN18 rgcn config (reference experiments/predict_links.py:119-195):
    # per-step sampling + negatives + edge dropout, encoder, DistMult decoder, BCE + L2 penalty, backward, optimiser
    'wn18_lp_step': dict(shape='wn18', kind='lp_step', in_f=200, out_f=200, decomp={'type': 'basis', 'num_bases': 2},
                         dtype='f32', vertical=False,
                         label='WN18-shaped rgcn training step (configs/rgcn/lp-WN18.yaml: 40,943 nodes, 18 relations, '
                               '141,442 training triples; 30,000 edge-neighbourhood positives + 10 negatives each, edge '
                               'dropout 0.5, 1 layer 200 -> 200 with basis B=2, DistMult, Adam): sampler + encoder + decoder '
                               '+ loss + backward + optimiser step'),
    'syn': dict(shape='syn', kind='nc', in_f=512, out_f=512, decomp={'type': 'block', 'num_blocks': 32}, dtype='bf16',
                vertical=True, raw=True,
                label='synthetic 5M-node / 256-rel / 200M-edge layer, block-diagonal nb=32, 512->512, bf16'),
    # the same graph with full (undecomposed) 512 x 512 weights per relation: 105 TFLOP per direction, a true GEMM with
    # gathered rows -> tcgen05 tensor-core kernel (propagate_umma.cuh)
    'syn_none': dict(shape='syn', kind='nc', in_f=512, out_f=512, decomp=None, dtype='bf16', vertical=True, raw=True,
                     label='synthetic 5M-node / 256-rel / 200M-edge layer, no decomposition (256 dense 512x512 weights), '
                           '512->512, bf16'),
}


def _synthetic():
    """torch_rgcn_b200/synthetic.py loaded by path: importing the package would dlopen librgcn_b200.so, which the
    reference arm must never map."""
    import importlib.util
    if 'rgcn_bench_synthetic' not in sys.modules:
        spec = importlib.util.spec_from_file_location('rgcn_bench_synthetic',
                                                      os.path.join(ROOT, 'torch_rgcn_b200', 'synthetic.py'))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        sys.modules['rgcn_bench_synthetic'] = mod
    return sys.modules['rgcn_bench_synthetic']


def load_reference():
    """The UNMODIFIED reference package, pip-installed from /root/reference into baseline/_ref by
    __graft_entry__.build() (git-ignored, travels to the GPU box).  None if that install is absent."""
    ref = os.path.join(ROOT, 'baseline', '_ref')
    if not os.path.isdir(os.path.join(ref, 'torch_rgcn')):
        return None
    if ref not in sys.path:
        sys.path.insert(0, ref)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        import torch_rgcn.layers as ref_layers
    return ref_layers


def dist_info():
    rank = int(os.environ.get('RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    return rank, world, local


# ------------------------------------------------------------------------------------------------------
# workload construction
# ------------------------------------------------------------------------------------------------------
def build_triples(wl, device, scale=1.0, seed=0, skew=False):
    syn = _synthetic()
    SHAPES, _rt = syn.SHAPES, syn.random_triples

    def random_triples(*a, **k):
        return _rt(*a, rel_dist='zipf' if skew else 'uniform', node_skew=skew, **k)
    N, R, E = SHAPES[wl['shape']]
    if scale != 1.0:
        N, E = max(int(N * scale), 64), max(int(E * scale), 64)
    if wl.get('raw'):                                # arbitrary triples over 2R relation ids, no inverse/self structure
        t = random_triples(N, 2 * R, 2 * E, seed=seed, device=device)
        return t, N, 2 * R, t.size(0)
    if wl['kind'] == 'lp':
        t = random_triples(N, R, E // 2, seed=seed, device=device)     # 50 % edge dropout of the train triples
        return t, N, 2 * R + 1, 3 * t.size(0) + N
    t = random_triples(N, R, E, seed=seed, device=device)
    return t, N, 2 * R + 1, 2 * E + N


def algorithmic_bytes(wl, N, Rp, nnz):
    """forward B_f and backward B_b in bytes (int32 indices, fp32 output/gradients, explicit val)."""
    I, O = wl['in_f'], wl['out_f']
    bx = 2 if wl['dtype'] == 'bf16' else 4
    d = wl['decomp'] or {}
    if d.get('type') == 'block':
        w = Rp * I * O // d['num_blocks']
    elif d.get('type') == 'basis':
        w = d['num_bases'] * I * O + Rp * d['num_bases']
    else:
        w = Rp * I * O
    b_f = nnz * (I * bx + 12) + 4 * (N + 1) + N * O * 4 + w * 4
    b_b = nnz * (O * 4 + 12) + nnz * (I * bx + 8) + N * O * 4 + N * I * 4 + 2 * w * 4
    return b_f, b_b, I * bx + 12


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        sm.sort()
        busy = [x for x in sm if mx and x > 0.5 * mx] or sm
        return {'sm_mhz': busy[len(busy) // 2] if busy else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    from torch_rgcn_b200 import _lib
    from torch_rgcn_b200.layers import RelationalGraphConvolutionNC, RelationalGraphConvolutionLP
    from torch_rgcn_b200.parallel import RelationShardedNC, RowShardedNC
    from torch_rgcn_b200.utils import add_inverse_and_self

    rank, world, local = dist_info()
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    wl = WORKLOADS[args.workload]
    t, N, Rp, nnz = build_triples(wl, dev, skew=args.skew)
    lp_graph, lp_graph_launches = False, 0
    xdt = torch.bfloat16 if wl['dtype'] == 'bf16' else torch.float32
    I, O = wl['in_f'], wl['out_f']
    torch.manual_seed(2)
    t_build = None
    if wl['kind'] == 'nc':
        tp = t if wl.get('raw') else add_inverse_and_self(t, N, (Rp - 1) // 2, device=dev)
        layer = RelationalGraphConvolutionNC(triples=tp, num_nodes=N, num_relations=Rp, in_features=I, out_features=O,
                                             decomposition=wl['decomp'], vertical_stacking=wl['vertical']).to(dev)
        if world > 1:
            layer = RowShardedNC(layer) if args.shard == 'rows' else RelationShardedNC(layer)
        call = lambda x: layer(x)                                            # noqa: E731
    else:
        layer = RelationalGraphConvolutionLP(num_nodes=N, num_relations=Rp, in_features=I, out_features=O,
                                             decomposition=wl['decomp'], vertical_stacking=wl['vertical'],
                                             b_init='zeros').to(dev).eval()
        layer.validate_triples = False          # no per-forward host sync (the reference's asserts do sync)
        call = lambda x: layer(t, x)                                         # noqa: E731
        if os.environ.get('RGCN_LP_GRAPH', '1') != '0':
            # the whole per-step graph build + propagation (and its backward) as ONE CUDA graph each
            try:
                from torch_rgcn_b200.layers import graph_lp_layer
                l0 = _lib.lib.rgcn_launch_count()
                graphed = graph_lp_layer(layer, t, torch.randn(N, I, device=dev, requires_grad=True))
                torch.cuda.synchronize()
                # make_graphed_callables runs 3 warm-up iterations and 1 capture of forward + backward: engine kernels
                # per replayed step (graph replays do not pass through the library's launch counter)
                lp_graph_launches = (_lib.lib.rgcn_launch_count() - l0) // 4
                call = lambda x: graphed(t, x)                               # noqa: E731
                lp_graph = True
            except Exception as exc:  # noqa: BLE001
                print(f'CUDA graph capture of the LP layer failed ({type(exc).__name__}: {str(exc)[:200]}); eager path',
                      file=sys.stderr, flush=True)
    gen = torch.Generator(device=dev).manual_seed(1)
    X = torch.randn(N, I, device=dev, generator=gen).to(xdt)
    G = torch.randn(N, O, device=dev, generator=gen)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # 256 MB > 126 MB L2

    def step_with(mod, x_in, g_in):
        x = x_in.detach().requires_grad_(True)
        out = mod(x)
        out.backward(g_in)
        return out.detach(), x.grad

    def step(x_in, g_in, ev=None):
        x = x_in.detach().requires_grad_(True)
        if ev:
            ev[0].record()
        out = call(x)
        if ev:
            ev[1].record()
        out.backward(g_in)
        if ev:
            ev[2].record()
        return out, x.grad

    sharded_parity = None
    if world > 1 and wl['kind'] == 'nc':
        # the sharded layer against the single-GPU engine on the same full-size inputs, on every rank, before timing
        torch.manual_seed(2)
        single = RelationalGraphConvolutionNC(triples=tp, num_nodes=N, num_relations=Rp, in_features=I, out_features=O,
                                              decomposition=wl['decomp'], vertical_stacking=wl['vertical']).to(dev)
        o_s, gx_s = step_with(single, X, G)
        gp_s = [p.grad.clone() for p in single.parameters()]
        inner = layer.layer
        for p in inner.parameters():
            p.grad = None
        o_m, gx_m = step(X, G)
        layer.sync_parameter_grads()
        tol = 3e-2 if wl['dtype'] == 'bf16' else 2e-4

        def rel_err(a, b):
            return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()
        errs = [rel_err(o_m, o_s), rel_err(gx_m, gx_s)] + [rel_err(p.grad, g) for p, g in zip(inner.parameters(), gp_s)]
        bad = torch.tensor([float(max(errs) > tol)], device=dev)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        sharded_parity = 'ok' if bad.item() == 0 else f'FAILED: max relative error {max(errs):.3g} on rank {rank} (tolerance {tol})'
        for p in inner.parameters():
            p.grad = None
        del single, o_s, gx_s, gp_s, o_m, gx_m
        torch.cuda.empty_cache()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    step(X, G)                                   # first call builds (and caches) the NC graph plan
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t0
    for _ in range(max(args.warmup - 1, 0)):
        flush.zero_()
        step(X, G)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timed region: exactly K steps, per-step events so the L2 flush is not counted
    sampler = ClockSampler(local)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    barrier()
    sampler.start()
    launches0 = _lib.lib.rgcn_launch_count()
    for k in range(args.steps):
        flush.zero_()
        step(X, G, evs[k])
    barrier()
    launches = _lib.lib.rgcn_launch_count() - launches0
    if lp_graph:
        launches = lp_graph_launches * args.steps            # kernel nodes of the replayed graphs
    t_fwd = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    t_bwd = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    tt = torch.tensor([t_fwd + t_bwd, t_fwd, t_bwd], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms_step, ms_fwd, ms_bwd = tt.tolist()

    # ---- end to end: host (pinned) buffers in and out, copies inside the timed region
    params = [p for p in layer.parameters()]
    # multi-GPU: a rank uploads / downloads only its 1/world slice of the rows over PCIe; the slices of X and G are
    # all-gathered over NVLink before the step (the host buffers of the ranks together hold each tensor once)
    rows_per = (N + world - 1) // world
    r_lo, r_hi = min(rank * rows_per, N), min((rank + 1) * rows_per, N)
    hX = X[r_lo:r_hi].cpu().pin_memory(); hG = G[r_lo:r_hi].cpu().pin_memory()
    hOut = torch.empty(r_hi - r_lo, O, dtype=torch.float32).pin_memory()
    hGX = torch.empty(r_hi - r_lo, I, dtype=xdt).pin_memory()
    hGP = [torch.empty_like(p, device='cpu').pin_memory() for p in params] if rank == 0 else []
    h2d = hX.numel() * hX.element_size() + hG.numel() * hG.element_size()
    d2h = hOut.numel() * 4 + hGX.numel() * hGX.element_size() + sum(h.numel() * 4 for h in hGP)
    io = torch.tensor([float(h2d), float(d2h)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(io, op=dist.ReduceOp.SUM)
    h2d, d2h = int(io[0].item()), int(io[1].item())            # whole-job bytes per step
    stageX = torch.zeros(rows_per, I, dtype=xdt, device=dev)
    stageG = torch.zeros(rows_per, O, dtype=torch.float32, device=dev)

    # Steps are pipelined the way a training loop would: copies run on their own streams, so the upload of step k+1
    # and the download of step k overlap compute (PCIe is full duplex); every byte still moves inside the timed region.
    s_h2d, s_d2h = torch.cuda.Stream(), torch.cuda.Stream()
    dX = [torch.empty(world * rows_per, I, dtype=xdt, device=dev) for _ in range(2)]
    dG = [torch.empty(world * rows_per, O, dtype=torch.float32, device=dev) for _ in range(2)]
    slot_free = [None, None]

    def e2e_step(k):
        slot = k % 2
        cur = torch.cuda.current_stream()
        with torch.cuda.stream(s_h2d):
            if slot_free[slot] is not None:
                s_h2d.wait_event(slot_free[slot])
            if world == 1:
                dX[slot].copy_(hX, non_blocking=True)
                dG[slot].copy_(hG, non_blocking=True)
            else:
                stageX[: r_hi - r_lo].copy_(hX, non_blocking=True)
                stageG[: r_hi - r_lo].copy_(hG, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(s_h2d)
        cur.wait_event(ready)
        if world > 1:
            dist.all_gather_into_tensor(dX[slot], stageX)
            dist.all_gather_into_tensor(dG[slot], stageG)
            staged = torch.cuda.Event()
            staged.record(cur)
            s_h2d.wait_event(staged)                           # the staging buffers may be refilled
        for p in params:
            p.grad = None
        x = dX[slot][:N].detach().requires_grad_(True)
        out = call(x)
        fwd_done = torch.cuda.Event()
        fwd_done.record(cur)
        out.backward(dG[slot][:N])
        bwd_done = torch.cuda.Event()
        bwd_done.record(cur)
        slot_free[slot] = bwd_done
        with torch.cuda.stream(s_d2h):
            s_d2h.wait_event(fwd_done)
            hOut.copy_(out.detach()[r_lo:r_hi], non_blocking=True)
            out.record_stream(s_d2h)
            s_d2h.wait_event(bwd_done)
            hGX.copy_(x.grad[r_lo:r_hi], non_blocking=True)
            x.grad.record_stream(s_d2h)
            for h, p in zip(hGP, params):
                if p.grad is not None:
                    h.copy_(p.grad, non_blocking=True)
                    p.grad.record_stream(s_d2h)

    e2e_step(0)
    torch.cuda.current_stream().wait_stream(s_d2h)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for k in range(args.steps):
        e2e_step(k + 1)
    torch.cuda.current_stream().wait_stream(s_d2h)
    e1.record()
    barrier()
    clocks = sampler.stop()
    te = torch.tensor([e0.elapsed_time(e1) / args.steps], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    ms_e2e = te.item()

    b_f, b_b, per_edge = algorithmic_bytes(wl, N, Rp, nnz)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except OSError:
        pass
    peak = peaks.get('hbm_gbs', 6650.0)
    peak_src = 'measured (MEASURED_PEAKS.json hbm_gbs)' if 'hbm_gbs' in peaks else 'fallback 6.65 TB/s'
    # per-rank bytes: a relation shard gathers nnz/world edges but still writes the full (N, O) partial output
    bf_rank = (b_f - N * O * 4) / world + N * O * 4
    bb_rank = (b_b - N * O * 4 - N * I * 4) / world + N * O * 4 + N * I * 4
    ach_f = bf_rank / (ms_fwd * 1e-3) / 1e9
    ach_b = bb_rank / (ms_bwd * 1e-3) / 1e9
    # tight byte model of what the backward kernels actually have to move (bf16 layers gather X[o] AND the bf16 copy
    # of grad_out[s] once per edge, 16 B of indices / weight per edge, one cast pass over grad_out, one write of the
    # feature gradient in the feature dtype) -- SURVEY's B_b assumes an fp32 grad_out gather plus a second X gather
    bx = 2 if wl['dtype'] == 'bf16' else 4
    w_bytes = (b_b - nnz * (O * 4 + 12) - nnz * (I * bx + 8) - N * O * 4 - N * I * 4) / 2
    bb_tight = nnz * (I * bx + O * bx + 16) / world + N * O * (4 + (bx if bx == 2 else 0)) + N * I * bx + 2 * w_bytes
    ach_b_tight = bb_tight / (ms_bwd * 1e-3) / 1e9
    ach_s = (bf_rank + bb_rank) / (ms_step * 1e-3) / 1e9
    # which kernels served the step: the fused row-block kernel (propagate_fused.cuh) or the two-phase kernels
    plan = None
    inner = getattr(layer, 'layer', layer)
    if getattr(inner, '_plan_cache', None):
        plan = inner._plan_cache[1]
    if getattr(layer, '_local', None) is not None:
        plan = layer._local
    if getattr(layer, '_plans', None) is not None:           # row-sharded: (key, forward plan, backward plan)
        plan = layer._plans[1]
    fused = bool(plan is not None and getattr(plan, 'fuse_rows', 0) > 0 and plan.fused_ok[0])
    fused_bwd = bool(fused and plan.fused_ok[1])
    traffic = {}
    try:       # DRAM bytes per launch from the committed `ncu --set full` capture of this workload (1 GPU)
        key = args.workload + ('_fused2' if fused_bwd else '_fused' if fused else '')
        traffic = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get(key, {}) if world == 1 else {}
    except OSError:
        pass
    fwd_kernel = ('forward edge gather: rgcn_forward = ONE fused row-block kernel, k_rowblock (TMA gather4 of the source '
                  'rows, per-relation MMA, degree normalisation and row sums in shared memory; no per-edge message '
                  'leaves the SM)') if fused else \
        'forward edge gather: rgcn_forward = gather+transform kernel + row-sum kernel'
    bwd_kernel = ('rgcn_backward = bf16 cast + bias grad, weight-gradient MMA pass, fused row-block kernel for the '
                  'feature gradient') if fused_bwd else \
        'rgcn_backward = bf16 cast + bias grad, fused feature/weight gradient kernel, row-sum'
    line = {
        'metric': 'rgcn_layer_edges_per_sec_fwd_bwd', 'value': nnz / (ms_step * 1e-3), 'unit': 'edges/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_step,
        'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None, 'dtype': wl['dtype'], 'data': 'synthetic',
        'config': {'workload': wl['label'] + (' [skewed: cubic node skew, Zipf relations]' if args.skew else ''),
                   'name': args.workload, 'num_nodes': N, 'num_relations': Rp, 'nnz': nnz,
                   'l2': 'L2 flushed (256 MB write) between timed steps; flush outside the event pairs',
                   'parallelism': ((f'row-sharded x{world}: every rank owns 1/{world} of the output rows; forward rows are '
                                    f'stored to all ranks by the fused kernel itself over NVLink peer memory, '
                                    f'feature-gradient rows pushed by the copy engines; parameter gradients all-reduced')
                                   if args.shard == 'rows' else
                                   f'relation-sharded x{world}, one all-reduce of out (fwd) and of grad_features (bwd)')
                   if world > 1 else 'single GPU',
                   'kernels': ('fused row-block TMA kernel for the forward' + (' and the feature gradient' if fused_bwd else
                                                                                  '; two-phase tensor-core kernels for the backward')
                               + f" (RGCN_FUSED={os.environ.get('RGCN_FUSED', '1')})") if fused else 'two-phase (messages through HBM)',
                   'graph_plan': 'built once at first call (outside timed region)' if wl['kind'] == 'nc'
                   else ('rebuilt every step (inside timed region)' +
                         ('; plan build + propagation replayed as one CUDA graph per direction' if lp_graph else ''))},
        'ms_fwd': ms_fwd, 'ms_bwd': ms_bwd, 'first_call_s_incl_plan_build': t_build,
        'e2e': {'value': nnz / (ms_e2e * 1e-3), 'unit': 'edges/s', 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': d2h},
        'gpu_launches': int(launches),
        'sharded_parity': sharded_parity,
        'roofline': {'kernel': fwd_kernel,
                     'bound': 'hbm', 'achieved': ach_f, 'peak': peak, 'unit': 'GB/s', 'frac': ach_f / peak,
                     'traffic': traffic.get('fwd_dram_bytes'), 'traffic_source': traffic.get('source'),
                     'peak_source': peak_src,
                     'algorithmic_bytes_per_launch': bf_rank, 'bytes_per_edge': per_edge},
        'roofline_bwd': {'kernel': bwd_kernel,
                         'bound': 'hbm', 'achieved': ach_b, 'peak': peak, 'unit': 'GB/s', 'frac': ach_b / peak,
                         'traffic': traffic.get('bwd_dram_bytes'), 'algorithmic_bytes_per_step': bb_rank,
                         'byte_model': 'SURVEY 8(d) B_b (fp32 grad_out gather + second X gather)',
                         'tight_model': {'bytes': bb_tight, 'achieved': ach_b_tight, 'frac': ach_b_tight / peak,
                                         'definition': 'one X[o] + one grad_out[s] row per edge in the feature dtype, '
                                                       '16 B/edge of indices and edge weight, cast pass over grad_out, '
                                                       'feature gradient written once in the feature dtype'}},
        'roofline_step': {'definition': 'SURVEY 8(d): (B_f + B_b) / (t_fwd + t_bwd)', 'achieved': ach_s, 'peak': peak,
                          'unit': 'GB/s', 'frac': ach_s / peak},
        'clocks': clocks,
    }
    if not wl['decomp'] and wl['dtype'] == 'bf16' and I % 64 == 0 and O % 64 == 0:
        # dense weights over bf16 features: per relation a GEMM with gathered rows (tcgen05 kernel k_gemm_umma); the
        # forward is bounded by the tensor cores, not by HBM
        tpeak = peaks.get('bf16_tflops_sustained', 1400.0)
        flops = 2.0 * nnz * I * O / world
        line['config']['kernels'] = ('tcgen05 gathered GEMM (k_gemm_umma: TMA gather4 -> UMMA 128 x 256 x 16, TMEM accumulators) '
                                     'for the forward and the feature gradient; fp32 register-tiled weight gradient')
        line['roofline_tensor'] = {'kernel': 'k_gemm_umma (forward)', 'bound': 'tensor', 'achieved': flops / (ms_fwd * 1e-3) / 1e12,
                                   'peak': tpeak, 'unit': 'TFLOP/s', 'frac': flops / (ms_fwd * 1e-3) / 1e12 / tpeak,
                                   'peak_source': 'MEASURED_PEAKS.json bf16_tflops_sustained (cuBLAS 8192^3)',
                                   'flops_per_launch': flops}
    syn = None
    if args.workload == 'am64' and not args.skew and not args.no_subrecords and os.environ.get('RGCN_BENCH_SYN', '1') != '0':
        # config 5 of BASELINE.json next to the headline, at every GPU count (all ranks take part)
        del layer, X, G, flush, dX, dG, stageX, stageG
        torch.cuda.empty_cache()
        syn = syn_record(rank, world, dev, args.shard)
        layer = X = G = flush = dX = dG = None
    if rank == 0:
        if syn is not None:
            line.setdefault('sub_records', {})['syn'] = syn
        if world == 1 and args.workload == 'am64' and not args.skew and not args.no_subrecords:
            # the two targets the north star quotes next to the headline: the fp32 (1e-4 parity) configuration of the
            # same graph, and the reference's own algorithm on this GPU at the largest scale that fits
            torch.cuda.empty_cache()
            sub = line.setdefault('sub_records', {})
            sub['am16_fp32'] = engine_layer_record('am16', 10, 3)
            ref_gpu = reference_gpu_record('am16', 0.25, 3, 1)
            sub['reference_gpu'] = ref_gpu
            if 'value' in ref_gpu:
                sub['am16_fp32_vs_reference_gpu_per_edge'] = sub['am16_fp32']['value'] / ref_gpu['value']
            if os.environ.get('RGCN_BENCH_SYN_NONE', '1') != '0':
                # the undecomposed 512 x 512 variant of the synthetic layer: the tcgen05 gathered-GEMM kernels
                try:
                    torch.cuda.empty_cache()
                    rec = engine_layer_record('syn_none', 2, 1)
                    flops = 2.0 * rec['nnz'] * 512 * 512
                    rec['tensor_tflops_fwd'] = flops / (rec['ms_fwd'] * 1e-3) / 1e12
                    rec['tensor_frac_fwd'] = rec['tensor_tflops_fwd'] / peaks.get('bf16_tflops_sustained', 1400.0)
                    rec['kernels'] = 'k_gemm_umma (forward, feature gradient), k_wgrad_umma: tcgen05 + TMA gather4 + TMEM'
                    sub['syn_none'] = rec
                except Exception as exc:  # noqa: BLE001
                    sub['syn_none'] = {'error': f'{type(exc).__name__}: {str(exc)[:200]}'}
        if world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_reference(args, budget_s=args.cpu_budget)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: CPU port of the reference's torch.sparse algorithm on a bounded sample
# ------------------------------------------------------------------------------------------------------
def reference_runner(wl, t, N, Rp, device, ref_layers):
    """(forward callable, parameter list, kind) for one layer call of the reference implementation on `device`:
    the unmodified reference layer class when baseline/_ref is installed, else the op-for-op port."""
    I, O = wl['in_f'], wl['out_f']
    torch.manual_seed(2)
    if ref_layers is not None:
        import warnings
        warnings.filterwarnings('ignore')
        if wl['kind'] == 'nc':
            if wl.get('raw'):
                tp = t
            else:
                from torch_rgcn.utils import add_inverse_and_self as ref_add
                tp = ref_add(t.to(device), N, (Rp - 1) // 2, str(torch.device(device).type))
            layer = ref_layers.RelationalGraphConvolutionNC(
                triples=tp.to(device), num_nodes=N, num_relations=Rp, in_features=I, out_features=O,
                decomposition=wl['decomp'], vertical_stacking=wl['vertical']).to(device)
            return (lambda x: layer(x)), list(layer.parameters()), 'reference'
        # the reference LP layer picks its device from torch.cuda.is_available() (layers.py:334, :461): for the CPU
        # arm on a GPU box, CUDA is hidden from it while it is constructed and called (the reference code is untouched)
        import contextlib

        @contextlib.contextmanager
        def visible_cuda():
            if torch.device(device).type != 'cpu':
                yield
                return
            saved = torch.cuda.is_available
            torch.cuda.is_available = lambda: False
            try:
                yield
            finally:
                torch.cuda.is_available = saved
        with visible_cuda():
            layer = ref_layers.RelationalGraphConvolutionLP(
                num_nodes=N, num_relations=Rp, in_features=I, out_features=O, decomposition=wl['decomp'],
                vertical_stacking=wl['vertical'], w_init='glorot-normal', b_init='zeros').to(device).eval()
        tt = t.to(device)

        def lp_fwd(x):
            with visible_cuda():
                return layer(tt, x)
        return lp_fwd, list(layer.parameters()), 'reference'
    from oracle import torch_sparse_port as port
    from oracle import rgcn_oracle as orc
    d = wl['decomp'] or {}
    params = {}
    if d.get('type') == 'block':
        nb = d['num_blocks']
        params['blocks'] = torch.randn(Rp - (1 if wl['kind'] == 'lp' else 0), nb, I // nb, O // nb, device=device)
        if wl['kind'] == 'lp':
            params['blocks_self'] = torch.randn(I, O, device=device)
    elif d.get('type') == 'basis':
        params['bases'] = torch.randn(d['num_bases'], I, O, device=device)
        params['comps'] = torch.randn(Rp, d['num_bases'], device=device)
    else:
        params['weights'] = torch.randn(Rp, I, O, device=device)
    params['bias'] = torch.zeros(O, device=device)
    for p in params.values():
        p.requires_grad_(True)
    if wl['kind'] == 'nc':
        tp = t if wl.get('raw') else torch.as_tensor(orc.add_inverse_and_self(t.cpu().numpy(), N, (Rp - 1) // 2))
        tp = tp.to(device)
        return (lambda x: port.nc_forward(tp, N, Rp, params, x, wl['vertical'])), list(params.values()), 'port'
    tt = t.to(device)
    return (lambda x: port.lp_forward(tt, N, Rp, params, x, wl['vertical'])), list(params.values()), 'port'


def cpu_reference(args, budget_s=20.0, steps=None, warmup=1):
    wl = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    N0, R0, E0 = _synthetic().SHAPES[wl['shape']]
    Rp = 2 * R0 if wl.get('raw') else 2 * R0 + 1
    # the reference materialises dense (R', N, d) fp32 temporaries; bound the largest to ~1 GB
    cap = 1e9 / (Rp * max(wl['in_f'], wl['out_f']) * 4)
    scale = min(1.0, cap / N0)
    t, N, Rp, nnz = build_triples(wl, 'cpu', scale=scale)
    I, O = wl['in_f'], wl['out_f']
    fwd, params, kind = reference_runner(wl, t, N, Rp, 'cpu', load_reference())
    X = torch.randn(N, I)
    G = torch.randn(N, O)

    def one():
        x = X.clone().requires_grad_(True)
        a = time.perf_counter()
        out = fwd(x)
        b = time.perf_counter()
        out.backward(G)
        c = time.perf_counter()
        for p in params:
            p.grad = None
        return b - a, c - b

    times = []
    for _ in range(warmup):
        one()
    n = steps if steps is not None else 3
    start = time.perf_counter()
    for k in range(n):
        times.append(one())
        if steps is None and time.perf_counter() - start > budget_s:
            break
    tf = sum(a for a, _ in times) / len(times)
    tb = sum(b for _, b in times) / len(times)
    what = ('the UNMODIFIED reference layer (baseline/_ref/torch_rgcn, pip-installed from /root/reference)'
            if kind == 'reference' else
            'oracle/torch_sparse_port.py (op-for-op CPU port of the reference torch.sparse path; baseline/_ref absent)')
    return {'value': nnz / (tf + tb), 'unit': 'edges/s', 'cores': cores, 'torch_threads': torch.get_num_threads(),
            'kind': kind, 's_fwd': tf, 's_bwd': tb, 'steps': len(times),
            'sample': f'{what} on a uniformly scaled graph: scale={scale:.4f}, N={N}, R\'={Rp}, nnz={nnz}, {I}->{O}, '
                      f'fp32, all {cores} host cores; per-edge rate is scale-invariant because the reference cost is '
                      f'O(R\'*N*d) with N/nnz fixed'}


def run_model(args):
    """Two-layer NodeClassifier training step on the B200 layers vs the CPU port of the reference wiring."""
    from torch_rgcn_b200 import _lib, models
    from oracle import torch_sparse_port as port
    wl = WORKLOADS[args.workload]
    dev = torch.device('cuda', 0)
    t, N, Rp, nnz = build_triples(wl, 'cpu')
    R = (Rp - 1) // 2
    torch.manual_seed(2)
    model = models.NodeClassifier(triples=t, nnodes=N, nrel=R, nhid=wl['hidden'], nlayers=2, nclass=wl['out_f'],
                                  decomposition=wl['decomp']).to(dev)
    g = torch.Generator().manual_seed(3)
    train_idx = torch.randperm(N, generator=g)[:max(N // 10, 8)].to(dev)
    labels = torch.randint(0, wl['out_f'], (train_idx.numel(),), generator=g).to(dev)
    crit = torch.nn.CrossEntropyLoss()

    def step():
        for p in model.parameters():
            p.grad = None
        loss = crit(model()[train_idx], labels)
        loss.backward()
        return loss

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    l0 = _lib.lib.rgcn_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    launches = _lib.lib.rgcn_launch_count() - l0
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step().item()                                   # e2e: loss read back every step (the model has no host inputs)
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    # CPU port of the same wiring (reference models.py:192-200)
    torch.set_num_threads(os.cpu_count() or 1)
    cpu = {n: p.detach().cpu().clone().requires_grad_(True) for n, p in model.named_parameters()}
    tp = model.triples_plus.cpu()
    ti, lb = train_idx.cpu(), labels.cpu()

    def cpu_step():
        p1 = {k[5:]: v for k, v in cpu.items() if k.startswith('rgc1.')}
        p2 = {k[5:]: v for k, v in cpu.items() if k.startswith('rgc2.')}
        h = torch.relu(port.nc_forward(tp, N, Rp, p1, None, False))
        out = port.nc_forward(tp, N, Rp, p2, h, True)
        crit(out[ti], lb).backward()
        for v in cpu.values():
            v.grad = None

    cpu_step()
    times = []
    for _ in range(3):
        a = time.perf_counter()
        cpu_step()
        times.append(time.perf_counter() - a)
    cpu_s = sum(times) / len(times)
    edges = 2 * nnz
    print(json.dumps({
        'metric': 'rgcn_layer_edges_per_sec_fwd_bwd', 'value': edges / (ms * 1e-3), 'unit': 'edges/s', 'n_gpus': 1,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': wl['dtype'], 'data': 'synthetic',
        'config': {'workload': wl['label'], 'name': args.workload, 'num_nodes': N, 'num_relations': Rp,
                   'edges_per_step': edges, 'l2': 'working set fits L2 (latency-bound shape); not flushed'},
        'e2e': {'value': edges / (ms_e2e * 1e-3), 'unit': 'edges/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches),
        'cpu_baseline': {'value': edges / cpu_s, 'unit': 'edges/s', 'cores': os.cpu_count(), 'kind': 'port',
                         's_per_step': cpu_s, 'sample': 'full-size model step with oracle/torch_sparse_port.py'}}),
        flush=True)


def run_decoder(args):
    """DistMult decoder step (reference predict_links.py:132-153 + models.py:239-244): negative sampling, scores of
    positives and negatives, Schlichtkrull L2 penalty, BCE-with-logits, backward to the node embeddings and the
    relation embeddings.  CPU baseline: the reference's own torch expressions on the host cores, bounded sample."""
    import torch.nn.functional as F
    from torch_rgcn_b200 import _lib
    from torch_rgcn_b200.decoder import DistMult, negative_sampling
    from torch_rgcn_b200.synthetic import SHAPES, random_triples
    wl = WORKLOADS[args.workload]
    dev = torch.device('cuda', 0)
    N, R, E = SHAPES[wl['shape']]
    d, rate = wl['in_f'], 10
    pos = random_triples(N, R, E, seed=0, device=dev)
    torch.manual_seed(2)
    dec = DistMult(R, d, N, R).to(dev)
    dec.validate_triples = False                       # no per-call host sync inside the timed region
    nodes = torch.randn(N, d, device=dev, requires_grad=True)
    labels = torch.cat([torch.ones(E, device=dev), torch.zeros(E * rate, device=dev)])

    def make_batch():
        neg = pos.clone()[:, None, :].expand(E, rate, 3).contiguous()
        return torch.cat([pos, negative_sampling(neg, N, 0.5, device=dev)], dim=0)

    def step(batch, ev=None):
        nodes.grad = None
        dec.relations.grad = None
        if ev:
            ev[0].record()
        scores = dec(batch, nodes)
        loss = F.binary_cross_entropy_with_logits(scores, labels) + 0.01 * dec.s_penalty(batch, nodes)
        if ev:
            ev[1].record()
        loss.backward()
        if ev:
            ev[2].record()
        return loss

    batch = make_batch()
    B = batch.size(0)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    for _ in range(max(args.warmup, 3)):
        step(batch)
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    l0 = _lib.lib.rgcn_launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()
        step(batch, evs[k])
    torch.cuda.synchronize()
    launches = _lib.lib.rgcn_launch_count() - l0
    ms_fwd = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    ms_bwd = sum(e[1].elapsed_time(e[2]) for e in evs) / args.steps
    # e2e: positives arrive from pinned host memory, negatives are sampled on the device, the loss goes back
    hpos = pos.cpu().pin_memory()
    dpos = torch.empty_like(pos)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        dpos.copy_(hpos, non_blocking=True)
        neg = dpos.clone()[:, None, :].expand(E, rate, 3).contiguous()
        step(torch.cat([dpos, negative_sampling(neg, N, 0.5, device=dev)], dim=0)).item()
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    ms = ms_fwd + ms_bwd
    # algorithmic bytes: forward reads the triple (24 B) and three embedding rows and writes a score; the penalty
    # re-reads the rows; backward re-reads them and adds two node rows (the relation gradient stays on chip)
    b_f = B * (24 + 3 * d * 4 + 4)
    b_b = B * (24 + 3 * d * 4 + 4 + 2 * d * 4)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except OSError:
        pass
    peak = peaks.get('hbm_gbs', 6650.0)
    line = {
        'metric': 'distmult_triples_per_sec_fwd_bwd', 'value': B / (ms * 1e-3), 'unit': 'triples/s', 'n_gpus': 1,
        'steps': args.steps, 'warmup': max(args.warmup, 3), 'ms_per_step': ms, 'higher_is_better': True,
        'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': wl['label'], 'name': args.workload, 'num_nodes': N, 'num_relations': R, 'dim': d,
                   'triples_per_step': B, 'l2': 'L2 flushed (256 MB write) between timed steps; the 21 MB node table is '
                                                'L2-resident within a step, so the gather model can exceed the HBM peak'},
        'ms_fwd': ms_fwd, 'ms_bwd': ms_bwd,
        'e2e': {'value': B / (ms_e2e * 1e-3), 'unit': 'triples/s', 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': hpos.numel() * 8, 'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches),
        'roofline': {'kernel': 'k_distmult_fwd + k_distmult_penalty (forward), k_distmult_bwd + k_distmult_penalty_bwd',
                     'bound': 'hbm', 'achieved': (b_f + b_b) / (ms * 1e-3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                     'frac': (b_f + b_b) / (ms * 1e-3) / 1e9 / peak, 'traffic': None,
                     'note': 'gathered rows come from L2 (node table 21 MB); fraction is of the HBM peak by convention'},
        'clocks': clocks,
    }
    if not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        nb = min(B, 200000)
        cb = batch[:nb].cpu()
        cn = nodes.detach().cpu().clone().requires_grad_(True)
        cr = dec.relations.detach().cpu().clone().requires_grad_(True)
        cl = labels[:nb].cpu()

        def cpu_step():                                 # reference layers.py:77-98 expressions
            s, p, o = cn[cb[:, 0]], cr[cb[:, 1]], cn[cb[:, 2]]
            scores = (s * p * o).sum(dim=-1)
            pen = s.pow(2).mean() + p.pow(2).mean() + o.pow(2).mean()
            (F.binary_cross_entropy_with_logits(scores, cl) + 0.01 * pen).backward()
            cn.grad = None
            cr.grad = None
        cpu_step()
        t0 = time.perf_counter()
        n = 0
        while n < 3 or time.perf_counter() - t0 < 5.0:
            cpu_step()
            n += 1
        cpu_s = (time.perf_counter() - t0) / n
        line['cpu_baseline'] = {'value': nb / cpu_s, 'unit': 'triples/s', 'cores': os.cpu_count(), 'kind': 'port',
                                's_per_step': cpu_s, 'steps': n,
                                'sample': f'first {nb} triples of the batch with the reference\'s torch expressions '
                                          f'(index, multiply, sum, mean, autograd) on the host cores'}
    print(json.dumps(line), flush=True)


def run_ranking(args):
    """Filtered ranking evaluation (reference utils/misc.py:60-110).  `value`: rank_triples for both sides with the node
    embeddings resident; e2e: evaluate() on a CompressionRelationPredictor — test triples from pinned host memory, the
    encoder once, both sides ranked, ranks and metrics back on the host.  CPU baseline: the reference's expressions
    (candidate triples (b, N, 3) -> DistMult scores -> filter -> rank) on a bounded sample of the test set."""
    from torch_rgcn_b200 import _lib
    from torch_rgcn_b200.evaluation import TrueTripleFilter, evaluate, rank_triples
    from torch_rgcn_b200.models import CompressionRelationPredictor
    from torch_rgcn_b200.synthetic import SHAPES, random_triples
    wl = WORKLOADS[args.workload]
    dev = torch.device('cuda', 0)
    N, R, E = SHAPES[wl['shape']]
    d, T = wl['in_f'], 5000
    train = random_triples(N, R, E, seed=0, device=dev)
    test = random_triples(N, R, T, seed=5, device=dev)
    known = torch.cat([train, random_triples(N, R, T, seed=6, device=dev), test], 0)
    enc = {'node_embedding': d, 'hidden1_size': 16, 'num_layers': 1, 'weight_init': 'glorot-normal', 'bias_init': 'zeros',
           'edge_dropout': {'general': 0.5, 'self_loop': 0.2, 'self_loop_type': 'schlichtkrull-dropout'}}
    decc = {'l2_penalty_type': 'schlichtkrull-l2', 'l2_penalty': 0.01, 'weight_init': 'standard-normal'}
    torch.manual_seed(2)
    model = CompressionRelationPredictor(nnodes=N, nrel=R, encoder_config=enc, decoder_config=decc).to(dev).eval()
    t0 = time.perf_counter()
    filt = TrueTripleFilter(known, N, R)
    torch.cuda.synchronize()
    filter_build_s = time.perf_counter() - t0
    with torch.no_grad():
        x = model.encode(train).contiguous()
    rel = model.scoring_function.relations.detach()

    def step():
        return rank_triples(test, x, rel, True, filt), rank_triples(test, x, rel, False, filt)

    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    W = max(args.warmup, 3)
    for _ in range(W):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    l0 = _lib.lib.rgcn_launch_count()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()
        evs[k][0].record()
        step()
        evs[k][1].record()
    torch.cuda.synchronize()
    launches = _lib.lib.rgcn_launch_count() - l0
    ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    htest = test.cpu().pin_memory()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n_e2e = max(3, min(args.steps, 10))
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n_e2e):
        mrr, hits, ranks = evaluate(model, train, htest.to(dev, non_blocking=True), filt, N, verbose=False)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_e2e = e0.elapsed_time(e1) / n_e2e
    flops = 2.0 * (2 * T) * N * d
    sm_mhz = clocks.get('sm_mhz') or clocks.get('sm_max_mhz') or 1965.0
    peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12                   # fp32 FMA lanes x 2 flop x clock
    line = {
        'metric': 'ranked_queries_per_sec', 'value': 2 * T / (ms * 1e-3), 'unit': 'queries/s', 'n_gpus': 1,
        'steps': args.steps, 'warmup': W, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': wl['label'], 'name': args.workload, 'num_nodes': N, 'num_relations': R, 'dim': d,
                   'test_triples': T, 'known_triples': int(known.size(0)), 'candidate_scores_per_step': 2 * T * N,
                   'l2': 'L2 flushed (256 MB write) between timed steps', 'filter_build_s': filter_build_s},
        'e2e': {'value': 2 * T / (ms_e2e * 1e-3), 'unit': 'queries/s', 'ms_per_step': ms_e2e,
                'h2d_bytes_per_step': htest.numel() * 8, 'd2h_bytes_per_step': 2 * T * 8,
                'note': 'evaluate(): H2D test triples, c-rgcn encoder once over the 141,442-triple graph, both sides '
                        'ranked, ranks to the host, MRR / hits@k on the host', 'mrr': mrr},
        'gpu_launches': int(launches),
        'roofline': {'kernel': 'k_rank_count', 'bound': 'fp32-fma', 'achieved': flops / (ms * 1e-3) / 1e12, 'peak': peak,
                     'unit': 'TFLOP/s', 'frac': flops / (ms * 1e-3) / 1e12 / peak, 'traffic': None,
                     'note': 'exact fp32 scores (ranks and ties must match the reference), so the candidate product '
                             'runs on the fp32 pipes, not the tensor cores; peak = 148 SMs x 128 lanes x 2 x SM clock; '
                             'achieved counts the whole step (query build, count, filter, finish)'},
        'clocks': clocks,
    }
    if not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        cx, cr, ct = x.cpu(), rel.cpu(), test.cpu()
        heads, tails = {}, {}
        for s_, p_, o_ in known.cpu().tolist():
            heads.setdefault((p_, o_), []).append(s_)
            tails.setdefault((s_, p_), []).append(o_)
        b = 16

        def cpu_batch(lo, head):                                  # utils/misc.py:75-101 expressions
            q = ct[lo:lo + b]
            n = q.size(0)
            cand = torch.arange(N)[None, :].expand(n, N)
            fixed = lambda c: q[:, c, None].expand(n, N)          # noqa: E731
            s_i, p_i, o_i = (cand, fixed(1), fixed(2)) if head else (fixed(0), fixed(1), cand)
            scores = (cx[s_i] * cr[p_i] * cx[o_i]).sum(-1)
            for i, (s_, p_, o_) in enumerate(q.tolist()):
                idx = [c for c in (heads[p_, o_] if head else tails[s_, p_]) if c != (s_ if head else o_)]
                scores[i, idx] = float('-inf')
            true = scores[torch.arange(n), q[:, 0 if head else 2]]
            raw = (scores > true[:, None]).sum(1)
            ties = (scores == true[:, None]).sum(1)
            return raw + (ties - 1) // 2 + 1
        cpu_batch(0, True)
        t0 = time.perf_counter()
        n = 0
        while n < 4 or time.perf_counter() - t0 < min(args.cpu_budget, 15.0):
            cpu_batch((n // 2) * b % (T - b), n % 2 == 0)
            n += 1
        cpu_s = (time.perf_counter() - t0) / n
        line['cpu_baseline'] = {'value': b / cpu_s, 'unit': 'queries/s', 'cores': os.cpu_count(), 'kind': 'port',
                                's_per_batch': cpu_s, 'batches': n,
                                'sample': f'{n} batches of {b} queries (the reference default batch size) with the '
                                          f"reference's expressions on the host cores; the reference also re-runs the "
                                          f'encoder per batch, which is NOT charged here'}
    print(json.dumps(line), flush=True)


def run_lp_step(args):
    """One epoch of experiments/predict_links.py:119-195 as ONE workload.  `value`: scored triples per second with the
    epoch inputs prefetched (StepInputPrefetcher, `depth` samplers in flight); `inline` sub-record: the same step with the
    sampler inside the step (depth 0, the reference's order of operations)."""
    import torch.nn.functional as F
    from torch_rgcn_b200 import _lib, models
    from torch_rgcn_b200.sampling import EdgeNeighborhoodSampler, StepInputPrefetcher, training_step_inputs
    syn = _synthetic()
    wl = WORKLOADS[args.workload]
    dev = torch.device('cuda', 0)
    N, R, E = syn.SHAPES[wl['shape']]
    S, rate, depth = 30000, 10, int(os.environ.get('RGCN_PREFETCH_DEPTH', '16'))
    train = syn.random_triples(N, R, E, seed=0, device=dev)
    torch.manual_seed(2)
    enc = {'node_embedding': 200, 'hidden1_size': 200, 'num_layers': 1, 'decomposition': dict(wl['decomp']),
           'edge_dropout': {'general': 0.5, 'self_loop': 0.2, 'self_loop_type': 'schlichtkrull-dropout'},
           # the shipped 'schlichtkrull-normal' raises inside the reference's (and the drop-in's) initialise_weights for
           # decomposed layers (layers.py:405-447 calls it without `shape`); glorot-normal has the same tensor shapes
           'weight_init': 'glorot-normal', 'include_gain': False, 'bias_init': 'zeros'}
    dec = {'l2_penalty_type': 'schlichtkrull-l2', 'l2_penalty': 0.01, 'weight_init': 'standard-normal', 'include_gain': False,
           'bias_init': 'zeros'}
    model = models.LinkPredictor(nnodes=N, nrel=R, encoder_config=enc, decoder_config=dec).to(dev).train()
    for m in model.modules():
        if hasattr(m, 'validate_triples'):
            m.validate_triples = False
    opt = torch.optim.Adam(model.parameters(), lr=0.01)
    sm = EdgeNeighborhoodSampler(train, N)
    kw = dict(graph_batch_size=S, neg_sample_rate=rate, edge_dropout_rate=0.5)

    def train_on(graph, batch, lbl):
        opt.zero_grad(set_to_none=True)
        scores, penalty = model(graph, batch)
        loss = F.binary_cross_entropy_with_logits(scores, lbl) + dec['l2_penalty'] * penalty
        loss.backward()
        opt.step()
        return loss

    def timed(fetch, steps, warmup):
        for _ in range(warmup):
            train_on(*fetch())
        torch.cuda.synchronize()
        l0 = _lib.lib.rgcn_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            loss = train_on(*fetch())
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps, _lib.lib.rgcn_launch_count() - l0, float(loss.item())

    W = max(args.warmup, 3)
    ms_inline, _, _ = timed(lambda: training_step_inputs(sm, N, **kw), min(args.steps, 10), W)
    pre = StepInputPrefetcher(sm, N, depth=depth, **kw)
    sampler = ClockSampler(0)
    sampler.start()
    ms, launches, loss = timed(pre.next, args.steps, W + depth)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        train_on(*pre.next()).item()                   # e2e: the loss is read back every epoch, like upstream's logging
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t0) * 1e3 / args.steps
    clocks = sampler.stop()
    sm.check()
    scored = S * (1 + rate)
    print(json.dumps({
        'metric': 'lp_training_step_scored_triples_per_sec', 'value': scored / (ms * 1e-3), 'unit': 'triples/s', 'n_gpus': 1,
        'steps': args.steps, 'warmup': W + depth, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': wl['label'], 'name': args.workload, 'num_nodes': N, 'positives': S, 'negatives': S * rate,
                   'prefetch_depth': depth,
                   'l2': 'working set of the step fits L2 (latency-bound shape); not flushed'},
        'inline_sampler': {'ms_per_step': ms_inline, 'note': 'sampler inside the step (reference order, no prefetch)'},
        'e2e': {'value': scored / (ms_e2e * 1e-3), 'unit': 'triples/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 4},
        'gpu_launches': int(launches), 'final_loss': loss, 'clocks': clocks}), flush=True)


def run_sampling(args):
    """Per-step graph construction (reference predict_links.py:123-148 with utils/misc.py:125-172).  `value`: picks per
    second of training_step_inputs with the training set resident; e2e: the same plus the D2H read of the sampled
    graph's size / a checksum (the training set is uploaded once per run, like upstream).  CPU baseline: the oracle's
    restatement of the reference's numpy loop on a bounded number of picks."""
    from torch_rgcn_b200 import _lib
    from torch_rgcn_b200.sampling import EdgeNeighborhoodSampler, training_step_inputs
    from torch_rgcn_b200.synthetic import SHAPES, random_triples
    wl = WORKLOADS[args.workload]
    dev = torch.device('cuda', 0)
    N, R, E = SHAPES[wl['shape']]
    S, rate = 30000, 10
    train = random_triples(N, R, E, seed=0, device=dev)
    t0 = time.perf_counter()
    sm = EdgeNeighborhoodSampler(train, N)
    torch.cuda.synchronize()
    build_s = time.perf_counter() - t0
    torch.manual_seed(1)

    def step():
        return training_step_inputs(sm, N, graph_batch_size=S, neg_sample_rate=rate, edge_dropout_rate=0.5)

    W = max(args.warmup, 3)
    for _ in range(W):
        step()
    sm.check()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    l0 = _lib.lib.rgcn_launch_count()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    for k in range(args.steps):
        flush.zero_()
        evs[k][0].record()
        step()
        evs[k][1].record()
    torch.cuda.synchronize()
    launches = _lib.lib.rgcn_launch_count() - l0
    ms = sum(e[0].elapsed_time(e[1]) for e in evs) / args.steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        graph, batch, lbl = step()
        int(graph.sum().item())
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    line = {
        'metric': 'sampled_positives_per_sec', 'value': S / (ms * 1e-3), 'unit': 'picks/s', 'n_gpus': 1,
        'steps': args.steps, 'warmup': W, 'ms_per_step': ms, 'higher_is_better': True, 'scaling': 'strong',
        'vs_baseline': None, 'dtype': 'i32', 'data': 'synthetic',
        'config': {'workload': wl['label'], 'name': args.workload, 'num_nodes': N, 'train_triples': E, 'sample_size': S,
                   'neg_sample_rate': rate, 'edge_dropout': 0.5, 'adjacency_build_s': build_s,
                   'l2': 'L2 flushed (256 MB write) between timed steps'},
        'e2e': {'value': S / (ms_e2e * 1e-3), 'unit': 'picks/s', 'ms_per_step': ms_e2e, 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 8, 'note': 'training set resident (uploaded once per run, as upstream); a '
                                                 'checksum of the sampled graph is read back every step'},
        'gpu_launches': int(launches),
        'roofline': {'kernel': 'k_sample_edge_neighborhood', 'bound': 'latency', 'achieved': None, 'peak': None,
                     'unit': 'picks/s', 'frac': None, 'traffic': None,
                     'note': 'a sequential process (pick i conditions on picks < i) run by one warp with its state in '
                             'shared memory; bounded by the dependent shared-memory / L2 latency chain per pick, not '
                             'by HBM or the tensor cores'},
        'clocks': clocks,
    }
    if not args.no_cpu_baseline:
        import numpy as np
        from oracle import sampling_oracle as so
        ct = train.cpu().numpy()
        picks = 1500
        t0 = time.perf_counter()
        so.edge_neighborhood(ct, N, picks, np.random.default_rng(0).random((picks, 2), dtype=np.float32))
        cpu_s = time.perf_counter() - t0
        line['cpu_baseline'] = {'value': picks / cpu_s, 'unit': 'picks/s', 'cores': 1, 'kind': 'port', 's_total': cpu_s,
                                'sample': f'{picks} picks of the numpy restatement of the reference loop (adjacency lists '
                                          f'rebuilt per call and an O(N) weight vector per pick, like upstream)'}
    print(json.dumps(line), flush=True)


def reference_gpu_record(workload, scale, steps, warmup):
    """The reference implementation on CUDA tensors of THIS box (the north star's '>= 1.0x the reference GPU
    torch.sparse path' comparison): the unmodified reference layer `.cuda()` when baseline/_ref is installed and its
    legacy sparse constructors still work on this torch, else the op-for-op port.  Halves the graph on OOM."""
    wl = WORKLOADS[workload]
    dev = torch.device('cuda', 0)
    ref_layers = load_reference()
    note = ''
    while True:
        try:
            t, N, Rp, nnz = build_triples(wl, dev, scale=scale)
            I, O = wl['in_f'], wl['out_f']
            try:
                fwd, params, kind = reference_runner(wl, t, N, Rp, dev, ref_layers)
                X = torch.randn(N, I, device=dev)
                fwd(X)                                   # the reference's legacy torch.cuda.sparse ctor (layers.py:277)
            except (torch.OutOfMemoryError, torch.AcceleratorError):
                raise
            except Exception as exc:  # noqa: BLE001
                if ref_layers is None:
                    raise
                note = f'unmodified reference failed on CUDA ({type(exc).__name__}: {str(exc)[:120]}); port used'
                ref_layers = None
                continue
            G = torch.randn(N, O, device=dev)
            times = []
            for k in range(warmup + steps):
                x = X.clone().requires_grad_(True)
                e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                e[0].record()
                out = fwd(x)
                e[1].record()
                out.backward(G)
                e[2].record()
                torch.cuda.synchronize()
                for p in params:
                    p.grad = None
                if k >= warmup:
                    times.append((e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])))
                del out, x
            break
        except (torch.OutOfMemoryError, torch.AcceleratorError) as exc:
            if isinstance(exc, torch.AcceleratorError):          # sticky error: the context is gone, report and stop
                return {'impl': 'reference-gpu', 'unavailable': f'{str(exc)[:160]} at scale {scale}'}
            torch.cuda.empty_cache()
            scale *= 0.5
            if scale < 1e-3:
                return {'impl': 'reference-gpu', 'unavailable': 'out of memory at every scale tried'}
    tf = sum(a for a, _ in times) / len(times)
    tb = sum(b for _, b in times) / len(times)
    rec = {'impl': 'reference-gpu', 'kind': kind, 'metric': 'rgcn_layer_edges_per_sec_fwd_bwd',
           'value': nnz / ((tf + tb) * 1e-3), 'unit': 'edges/s', 'ms_fwd': tf, 'ms_bwd': tb, 'steps': len(times),
           'config': {'workload': wl['label'], 'name': workload, 'scale': scale, 'num_nodes': N, 'nnz': nnz},
           'note': ('the unmodified reference layer (baseline/_ref) moved to CUDA' if kind == 'reference' else
                    'reference torch.sparse algorithm (oracle/torch_sparse_port.py) on CUDA tensors') +
                   ', fp32, CUDA-event timed, same box' + (('; ' + note) if note else ''),
           'peak_mem_gb': torch.cuda.max_memory_allocated() / 1e9}
    torch.cuda.empty_cache()
    return rec


def run_reference_gpu(args):
    print(json.dumps(reference_gpu_record(args.workload, args.ref_scale, args.steps, args.warmup)), flush=True)


def syn_record(rank, world, dev, shard_mode, steps=3, warmup=1):
    """BASELINE configs[4] (synthetic 5M-node / 256-rel / 200M-edge layer, 512 -> 512, nb=32, bf16) as a sub-record of
    every --gpus N line: device-resident fwd / bwd of the (row- or relation-) sharded layer, max over ranks."""
    import torch.distributed as dist
    from torch_rgcn_b200.layers import RelationalGraphConvolutionNC
    from torch_rgcn_b200.parallel import RelationShardedNC, RowShardedNC
    wl = WORKLOADS['syn']
    ok = torch.ones(1, device=dev)
    rec = {'workload': wl['label'], 'name': 'syn', 'n_gpus': world, 'steps': steps, 'warmup': warmup}
    try:
        t, N, Rp, nnz = build_triples(wl, dev)
        torch.manual_seed(2)
        layer = RelationalGraphConvolutionNC(triples=t, num_nodes=N, num_relations=Rp, in_features=wl['in_f'],
                                             out_features=wl['out_f'], decomposition=wl['decomp'],
                                             vertical_stacking=wl['vertical']).to(dev)
        if world > 1:
            layer = RowShardedNC(layer) if shard_mode == 'rows' else RelationShardedNC(layer)
        gen = torch.Generator(device=dev).manual_seed(1)
        X = torch.randn(N, wl['in_f'], device=dev, generator=gen).to(torch.bfloat16)
        G = torch.randn(N, wl['out_f'], device=dev, generator=gen)
    except Exception as exc:  # noqa: BLE001  (out of memory while building: report, keep the headline line)
        ok.zero_()
        rec['error'] = f'{type(exc).__name__}: {str(exc)[:200]}'
    if world > 1:
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    if ok.item() == 0:
        rec.setdefault('error', 'another rank failed while building the workload')
        return rec
    tf = tb = 0.0
    for k in range(warmup + steps):
        x = X.detach().requires_grad_(True)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e[0].record()
        out = layer(x)
        e[1].record()
        out.backward(G)
        e[2].record()
        torch.cuda.synchronize()
        if k >= warmup:
            tf += e[0].elapsed_time(e[1]) / steps
            tb += e[1].elapsed_time(e[2]) / steps
        del out, x
    tt = torch.tensor([tf + tb, tf, tb], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms, tf, tb = tt.tolist()
    b_f, b_b, _ = algorithmic_bytes(wl, N, Rp, nnz)
    rec.update({'nnz': nnz, 'ms_fwd': tf, 'ms_bwd': tb, 'value': nnz / (ms * 1e-3), 'unit': 'edges/s', 'scaling': 'strong',
                'gather_model_gbs_step_per_gpu': (b_f + b_b) / world / (ms * 1e-3) / 1e9,
                'note': 'X (5.1 GB) is far larger than L2; inputs larger than L2, no flush between steps'})
    del layer, X, G
    torch.cuda.empty_cache()
    return rec


def engine_layer_record(workload, steps, warmup):
    """Device-resident fwd / bwd times of this engine on one NC workload (sub-records of the headline line)."""
    from torch_rgcn_b200.layers import RelationalGraphConvolutionNC
    from torch_rgcn_b200.utils import add_inverse_and_self
    wl = WORKLOADS[workload]
    dev = torch.device('cuda', 0)
    t, N, Rp, nnz = build_triples(wl, dev)
    xdt = torch.bfloat16 if wl['dtype'] == 'bf16' else torch.float32
    torch.manual_seed(2)
    tp = t if wl.get('raw') else add_inverse_and_self(t, N, (Rp - 1) // 2, device=dev)
    layer = RelationalGraphConvolutionNC(triples=tp, num_nodes=N, num_relations=Rp, in_features=wl['in_f'],
                                         out_features=wl['out_f'], decomposition=wl['decomp'],
                                         vertical_stacking=wl['vertical']).to(dev)
    gen = torch.Generator(device=dev).manual_seed(1)
    X = torch.randn(N, wl['in_f'], device=dev, generator=gen).to(xdt)
    G = torch.randn(N, wl['out_f'], device=dev, generator=gen)
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
    tf = tb = 0.0
    for k in range(warmup + steps):
        flush.zero_()
        x = X.detach().requires_grad_(True)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        out = layer(x)
        e[1].record()
        out.backward(G)
        e[2].record()
        torch.cuda.synchronize()
        if k >= warmup:
            tf += e[0].elapsed_time(e[1]) / steps
            tb += e[1].elapsed_time(e[2]) / steps
    b_f, b_b, _ = algorithmic_bytes(wl, N, Rp, nnz)
    del layer, X, G, flush
    torch.cuda.empty_cache()
    return {'workload': wl['label'], 'name': workload, 'dtype': wl['dtype'], 'nnz': nnz, 'ms_fwd': tf, 'ms_bwd': tb,
            'value': nnz / ((tf + tb) * 1e-3), 'unit': 'edges/s',
            'gather_model_gbs_fwd': b_f / (tf * 1e-3) / 1e9,
            'note': 'parity bar 1e-4 vs the reference (tests/test_gpu_parity.py); X fits L2, so the gather model is '
                    'not an HBM fraction (SURVEY 8d)' if workload == 'am16' else ''}


def run_reference(args):
    rank, world, _ = dist_info()
    if rank != 0:
        return
    wl = WORKLOADS[args.workload]
    base = cpu_reference(args, steps=args.steps, warmup=args.warmup)
    ms = (base['s_fwd'] + base['s_bwd']) * 1e3
    line = {'impl': 'reference', 'metric': 'rgcn_layer_edges_per_sec_fwd_bwd', 'value': base['value'], 'unit': 'edges/s',
            'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': {'workload': wl['label'], 'name': args.workload},
            'cpu_baseline': base,
            'e2e': {'value': base['value'], 'unit': 'edges/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--workload', default='am64', choices=sorted(WORKLOADS))
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference', 'reference-gpu'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-subrecords', action='store_true', help='skip the am16_fp32 / reference_gpu sub-records')
    ap.add_argument('--ref-scale', type=float, default=1.0, help='graph scale for --impl reference-gpu')
    ap.add_argument('--skew', action='store_true', help='power-law node degrees + Zipf relation sizes (hub rows)')
    ap.add_argument('--cpu-budget', type=float, default=20.0)
    ap.add_argument('--shard', default='rows', choices=['relations', 'rows'],
                    help='multi-GPU partition of an NC layer: rows (default: every rank owns output rows, which the fused '
                         'kernel stores to all ranks over NVLink -- half the bytes of an all-reduce, see DESIGN.md 5) or '
                         'relations + all-reduce (the first design of the north star, kept for comparison)')
    args = ap.parse_args()
    if args.impl == 'reference-gpu':
        run_reference_gpu(args)
    elif args.impl == 'reference':
        run_reference(args)
    elif WORKLOADS[args.workload]['kind'] == 'nc_model':
        run_model(args)
    elif WORKLOADS[args.workload]['kind'] == 'decoder':
        run_decoder(args)
    elif WORKLOADS[args.workload]['kind'] == 'ranking':
        run_ranking(args)
    elif WORKLOADS[args.workload]['kind'] == 'sampling':
        run_sampling(args)
    elif WORKLOADS[args.workload]['kind'] == 'lp_step':
        run_lp_step(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == '__main__':
    main()
