"""MIPEncoder / SolutionEncoder containers (/root/reference/model/encoder.py)."""
import torch

from model.modules import GNNModel, ResidualAttentionBlock


class MIPEncoder(torch.nn.Module):
    """Instance encoder: bipartite GNN -> per-variable 128-d embedding (reference: encoder.py:7-22).
    The reference's ``mean_aggr`` attribute (a PyG aggregation that is never called) has no
    parameters and is not reproduced."""

    def __init__(self, emb_size=64, cons_nfeats=5, edge_nfeats=1, var_nfeats=17, gcn_mlp_layer_num=2):
        super().__init__()
        self.gnn_model = GNNModel(emb_size, cons_nfeats, edge_nfeats, var_nfeats, gcn_mlp_layer_num=gcn_mlp_layer_num)

    def forward(self, constraint_features, edge_indices, edge_features, variable_features):
        return self.gnn_model(constraint_features, edge_indices, edge_features, variable_features)


class SolutionEncoder(torch.nn.Module):
    """Parameter container for checkpoint compatibility (reference: encoder.py:24-36).  Encoding
    solutions is part of contrastive training, not of reverse sampling: out of scope."""

    def __init__(self, emb_num: int, attn_dim: int, num_heads: int, layers: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.embedding = torch.nn.Embedding(emb_num, attn_dim)
        self.width = attn_dim
        self.layers = layers
        self.resblocks = torch.nn.ModuleList([ResidualAttentionBlock(attn_dim, num_heads, attn_mask) for _ in range(layers)])

    def forward(self, x, key_padding_mask=None):
        raise NotImplementedError("SolutionEncoder.forward belongs to CMSP training (out of scope of the sampling path)")
