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
    """Token embedding of a (padded) 0/1/2 solution + attention blocks (reference: encoder.py:24-36), forward only: the x-side
    features of CMSP.get_features(mip, x) -- the latents the diffusion model is trained to reproduce (train_diffusion.py).  The
    reference adds the block's output to its input although the block is residual already (`x = x + module(x, mask)`): kept."""

    def __init__(self, emb_num: int, attn_dim: int, num_heads: int, layers: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.embedding = torch.nn.Embedding(emb_num, attn_dim)
        self.width = attn_dim
        self.layers = layers
        self.resblocks = torch.nn.ModuleList([ResidualAttentionBlock(attn_dim, num_heads, attn_mask) for _ in range(layers)])

    @torch.no_grad()
    def forward(self, x, key_padding_mask=None):
        from dip_b200 import ops
        if x.dim() != 2:
            raise ValueError("solutions must be (G, L) integer tokens")
        G, L = x.shape
        emb = self.embedding.weight.detach().contiguous()
        tok = x.to(torch.int64)
        if int(tok.min()) < 0 or int(tok.max()) >= emb.shape[0]:
            raise IndexError("solution token outside the embedding table")
        h = torch.stack([ops.embedding_rows(emb, tok[g].contiguous()) for g in range(G)])          # (G, L, 128)
        for module in self.resblocks:
            h = ops.axpby(h, module(h, key_padding_mask))                                        # x + module(x): encoder.py:35
        return h
