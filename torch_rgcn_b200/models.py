"""Node-classification models wired on the B200 layers (SURVEY §8 f, rank 1).

`NodeClassifier` and `EmbeddingNodeClassifier` keep the constructor arguments, buffers, sub-module and parameter
names of reference torch_rgcn/models.py:137-200 and :248-296, so checkpoints and the experiment scripts' L2 penalties
(`model.rgc1.weights` ...) carry over.  The link-prediction models of the reference are broken at this commit
 and are not mirrored; `RelationalGraphConvolutionLP` itself is in layers.py.
"""
import torch
import torch.nn.functional as F
from torch import nn

from .layers import RelationalGraphConvolutionNC
from .utils import add_inverse_and_self


class NodeClassifier(nn.Module):
    """Two-layer (or one-layer) R-GCN node classifier: featureless layer 1 (horizontal), ReLU, layer 2 (vertical)."""

    def __init__(self, triples=None, nnodes=None, nrel=None, nfeat=None, nhid=16, nlayers=2, nclass=None,
                 edge_dropout=None, decomposition=None, nemb=None):
        super().__init__()
        self.nlayers = nlayers
        assert (triples is not None or nnodes is not None or nrel is not None or nclass is not None), \
            "The following must be specified: triples, number of nodes, number of relations and number of classes!"
        assert 0 < nlayers < 3, "Only supports the following number of RGCN layers: 1 and 2."
        if nlayers == 1:
            nhid = nclass
        if nlayers == 2:
            assert nhid is not None, "Number of hidden layers not specified!"
        triples = torch.as_tensor(triples, dtype=torch.long)
        with torch.no_grad():
            self.register_buffer('triples', triples)
            self.register_buffer('triples_plus', add_inverse_and_self(triples, nnodes, nrel))
        self.rgc1 = RelationalGraphConvolutionNC(triples=self.triples_plus, num_nodes=nnodes, num_relations=nrel * 2 + 1,
                                                 in_features=nfeat, out_features=nhid, edge_dropout=edge_dropout,
                                                 decomposition=decomposition, vertical_stacking=False)
        if nlayers == 2:
            self.rgc2 = RelationalGraphConvolutionNC(triples=self.triples_plus, num_nodes=nnodes,
                                                     num_relations=nrel * 2 + 1, in_features=nhid, out_features=nclass,
                                                     edge_dropout=edge_dropout, decomposition=decomposition,
                                                     vertical_stacking=True)

    def forward(self):
        x = self.rgc1()
        if self.nlayers == 2:
            x = F.relu(x)
            x = self.rgc2(features=x)
        return x


class EmbeddingNodeClassifier(NodeClassifier):
    """e-rgcn: learned node embeddings -> diagonal-weight layer -> ReLU -> R-GCN layer (reference models.py:248-296)."""

    def __init__(self, triples=None, nnodes=None, nrel=None, nfeat=None, nhid=16, nlayers=2, nclass=None,
                 edge_dropout=None, decomposition=None, nemb=None):
        assert nemb is not None, "Size of node embedding not specified!"
        nfeat = nemb
        assert nlayers == 2, "For this model only 2 layers are normally configured (for now)"
        nhid = nemb
        super().__init__(triples, nnodes, nrel, nfeat, nhid, 1, nclass, edge_dropout, decomposition)
        self.rgcn_no_hidden = RelationalGraphConvolutionNC(triples=self.triples_plus, num_nodes=nnodes,
                                                           num_relations=nrel * 2 + 1, in_features=nfeat,
                                                           out_features=nhid, edge_dropout=edge_dropout,
                                                           decomposition=decomposition, vertical_stacking=False,
                                                           diag_weight_matrix=True)
        self.node_embeddings = nn.Parameter(torch.FloatTensor(nnodes, nemb))
        nn.init.kaiming_normal_(self.node_embeddings, mode='fan_in')

    def forward(self):
        x = self.rgcn_no_hidden(self.node_embeddings)
        x = F.relu(x)
        return self.rgc1(features=x)
