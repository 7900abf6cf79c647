"""Per-step graph construction for link prediction on the device.

Device counterparts of reference utils/misc.py:112-172 (`select_sampling`, `uniform_sampling`, `edge_neighborhood`)
and of the general edge dropout in experiments/predict_links.py:143-148, plus `training_step_inputs`, which assembles
what one epoch of predict_links.py:123-148 builds (positives, negatives, labels, message-passing graph).  Everything
returns CUDA tensors; there is no CPU fallback (torch_rgcn_b200/csrc/sampling.cu).

The reference samples with numpy's / Python's global generators.  Here the randomness is torch's CUDA generator
(`torch.rand`, `torch.randperm` on the device), so runs are reproducible under `torch.manual_seed` but do not replay the
reference's draws; the distribution is the reference's (checked statistically against the reference's own samples and
pick for pick against a CPU restatement on shared uniforms, tests/test_gpu_sampling.py).
"""
import torch

from . import _lib
from .decoder import negative_sampling


def _as_triples(triples, device=None):
    t = torch.as_tensor(triples, dtype=torch.long)
    if device is None:
        _lib.require_cuda()
        device = t.device if t.is_cuda else torch.device('cuda', torch.cuda.current_device())
    return t.to(device).reshape(-1, 3).contiguous()


def take_triples(triples, index):
    """triples[index] through the engine (index int32 or int64, on the device); raises IndexError on a bad index."""
    _lib.require_cuda(triples, index)
    assert index.dtype in (torch.int32, torch.int64)
    index = index.contiguous()
    out = torch.empty(index.numel(), 3, dtype=torch.long, device=triples.device)
    status = torch.zeros(1, dtype=torch.int32, device=triples.device)
    with torch.cuda.device(triples.device):
        _lib.check(_lib.lib.rgcn_take_triples(_lib.ptr(triples), triples.size(0), _lib.ptr(index),
                                              1 if index.dtype == torch.int64 else 0, index.numel(), _lib.ptr(out),
                                              _lib.ptr(status), _lib.stream_ptr()))
    return out, status


class EdgeNeighborhoodSampler:
    """Edge-neighbourhood sampling (reference utils/misc.py:125-172) over a fixed training set.

    The vertex adjacency is built once (the reference rebuilds Python lists on every call); `sample` then runs the
    sequential frontier process in one on-chip kernel."""

    def __init__(self, train_triples, num_nodes, device=None):
        t = _as_triples(train_triples, device)
        self.triples, self.num_nodes, self.num_edges, self.device = t, int(num_nodes), t.size(0), t.device
        E, dev = self.num_edges, t.device
        self.adj_ptr = torch.empty(self.num_nodes + 1, dtype=torch.int32, device=dev)
        self.adj = torch.empty(max(2 * E, 1), 2, dtype=torch.int32, device=dev)       # {edge index, other end} per entry
        self.adj_edge, self.adj_other = self.adj[:, 0], self.adj[:, 1]
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        ws_bytes = _lib.lib.rgcn_sampler_build_workspace_bytes(E)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.rgcn_sampler_build(_lib.ptr(t), E, self.num_nodes, _lib.ptr(self.adj_ptr),
                                                   _lib.ptr(self.adj), _lib.ptr(status), _lib.ptr(ws), ws_bytes,
                                                   _lib.stream_ptr()))
        bad = int(status.item())
        assert bad == 0, f'{bad} training triples index a node >= {num_nodes}'
        self._ws_bytes = _lib.lib.rgcn_sample_workspace_bytes(E, self.num_nodes)

    def sample_indices(self, sample_size, uniforms=None):
        """Edge numbers of the sample in pick order, int32 (sample_size,).  `uniforms` (sample_size, 2) fp32 in [0, 1)
        replaces the generator draw (tests)."""
        S, dev = int(sample_size), self.device
        if S > self.num_edges:
            raise ValueError(f'sample_size {S} exceeds the {self.num_edges} training triples')
        if uniforms is None:
            uniforms = torch.rand(S, 2, device=dev)
        u = uniforms.to(device=dev, dtype=torch.float32).contiguous()
        assert u.numel() == 2 * S
        out = torch.empty(S, dtype=torch.int32, device=dev)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        ws = torch.empty(self._ws_bytes, dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.check(_lib.lib.rgcn_sample_edge_neighborhood(_lib.ptr(self.adj_ptr), _lib.ptr(self.adj),
                                                              self.num_edges, self.num_nodes,
                                                              _lib.ptr(u), S, _lib.ptr(out), _lib.ptr(status),
                                                              _lib.ptr(ws), self._ws_bytes, _lib.stream_ptr()))
        self._last_status = status                      # checked lazily: no host sync on the training path
        return out

    def sample(self, sample_size, uniforms=None):
        """(sample_size, 3) int64 triples on the device."""
        idx = self.sample_indices(sample_size, uniforms)
        out, _ = take_triples(self.triples, idx)
        return out

    def check(self):
        """Host-synchronising consistency check of the last sample (0 = fine)."""
        bad = int(self._last_status.item())
        if bad:
            raise RuntimeError(f'edge-neighbourhood sampler failed (status {bad})')


_sampler_keepalive = {}


def edge_neighborhood(train_triples, sample_size=30000, entities=None):
    """Drop-in for reference utils/misc.py:125 (same arguments; `entities` is the node dictionary, only its length is
    used, like upstream).  Returns a (sample_size, 3) CUDA tensor; `torch.tensor(result, device=...)` in the caller
    (predict_links.py:131) keeps working.  The adjacency is cached per training set object."""
    key = id(train_triples)
    hit = _sampler_keepalive.get(key)
    if hit is None or hit[0] is not train_triples:
        num_nodes = len(entities) if entities is not None else int(_as_triples(train_triples)[:, [0, 2]].max().item()) + 1
        _sampler_keepalive.clear()                      # one training set at a time
        hit = (train_triples, EdgeNeighborhoodSampler(train_triples, num_nodes))
        _sampler_keepalive[key] = hit
    return hit[1].sample(sample_size)


def uniform_sampling(graph, sample_size=30000, entities=None, train_triplets=None):
    """Drop-in for reference utils/misc.py:121-123 (`random.sample(graph, sample_size)`): sample_size distinct rows in
    random order."""
    t = _as_triples(graph)
    if sample_size > t.size(0):
        raise ValueError('Sample larger than population or is negative')      # random.sample's error
    idx = torch.randperm(t.size(0), device=t.device)[:sample_size]
    out, _ = take_triples(t, idx)
    return out


def select_sampling(method):
    """reference utils/misc.py:112-119"""
    method = method.lower()
    if method == 'uniform':
        return uniform_sampling
    if method == 'edge-neighborhood':
        return edge_neighborhood
    raise NotImplementedError(f'{method} sampling method has not been implemented!')


def edge_dropout(graph, edge_dropout_rate, perm=None):
    """General edge dropout of predict_links.py:143-148, literally: shuffle, then drop the FIRST round(keep_prob * n)
    rows (so a fraction `keep_prob` is dropped, not kept — upstream's arithmetic, identical at the shipped rate 0.5).
    `perm` replaces the generator draw."""
    t = _as_triples(graph)
    if not edge_dropout_rate > 0.0:
        return t
    n = t.size(0)
    keep_prob = 1 - edge_dropout_rate
    if perm is None:
        perm = torch.randperm(n, device=t.device)
    cut = round(keep_prob * n)
    out, _ = take_triples(t, perm.to(t.device)[cut:])
    return out


def training_step_inputs(sampler_or_train, num_nodes, graph_batch_size=None, neg_sample_rate=10, head_corrupt_prob=0.5,
                         edge_dropout_rate=0.0, training=True):
    """What predict_links.py:123-148 builds for one epoch: (graph, batch_idx, train_lbl), all on the device.

    `sampler_or_train`: an EdgeNeighborhoodSampler (edge-neighbourhood sampling of graph_batch_size positives) or a
    (E, 3) tensor (graph_batch_size None: the whole training set, as upstream; else uniform sampling)."""
    with torch.no_grad():
        if isinstance(sampler_or_train, EdgeNeighborhoodSampler):
            positives = sampler_or_train.sample(graph_batch_size)
        elif graph_batch_size is None:
            positives = _as_triples(sampler_or_train)
        else:
            positives = uniform_sampling(sampler_or_train, graph_batch_size)
        dev, B = positives.device, positives.size(0)
        negatives = positives.clone()[:, None, :].expand(B, neg_sample_rate, 3).contiguous()
        negatives = negative_sampling(negatives, num_nodes, head_corrupt_prob, device=dev)
        batch_idx = torch.cat([positives, negatives], dim=0)
        train_lbl = torch.cat([torch.ones(B, device=dev), torch.zeros(B * neg_sample_rate, device=dev)])
        graph = edge_dropout(positives, edge_dropout_rate) if training and edge_dropout_rate > 0.0 else positives
    return graph, batch_idx, train_lbl


class StepInputPrefetcher:
    """Keeps `depth` epochs' worth of `training_step_inputs` in flight, each on its own CUDA stream.

    Edge-neighbourhood sampling is one warp's sequential process (~43 ms per 30,000 picks at WN18) but it depends only
    on the training set and the generator, never on the model: the samples of the NEXT epochs can be drawn while this
    epoch trains.  Each sampler occupies one SM; `depth` of them run concurrently with the layer / decoder kernels of
    the current step on the remaining SMs, so the sampler's latency is paid `1 / depth` times per epoch.  The draws
    come from torch's CUDA generator in launch order, so a seeded run is reproducible.  (The reference samples on the
    host inside the epoch loop, experiments/predict_links.py:123-131.)"""

    def __init__(self, sampler_or_train, num_nodes, depth=8, **step_kwargs):
        import collections
        self.src, self.num_nodes, self.kwargs = sampler_or_train, num_nodes, step_kwargs
        self.streams = [torch.cuda.Stream() for _ in range(max(1, int(depth)))]
        self.queue = collections.deque()
        for st in self.streams:
            self._launch(st)

    def _launch(self, stream):
        stream.wait_stream(torch.cuda.current_stream())          # the training set / earlier frees are ready
        with torch.cuda.stream(stream):
            out = training_step_inputs(self.src, self.num_nodes, **self.kwargs)
            done = torch.cuda.Event()
            done.record(stream)
        self.queue.append((out, done, stream))

    def next(self):
        """(graph, batch_idx, train_lbl) of the oldest epoch in flight; starts drawing a new one."""
        out, done, stream = self.queue.popleft()
        cur = torch.cuda.current_stream()
        cur.wait_event(done)
        for t in out:
            t.record_stream(cur)
        self._launch(stream)
        return out
