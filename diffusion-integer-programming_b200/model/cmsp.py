"""CMSP instance encoder, inference half (/root/reference/model/cmsp.py:7-63)."""
import numpy as np
import torch

from dip_b200 import ops
from model.encoder import MIPEncoder, SolutionEncoder


class CMSP(torch.nn.Module):
    """Contrastive MIP-Solution Pretrain model; only ``get_features`` / ``encode_mip`` are on the
    sampling path.  Same constructor and state_dict layout as the reference (cmsp.py:11-23)."""

    def __init__(self, emb_num=3, emb_dim=128, n_heads=2, n_layers=2, padding_len=2000, position_emb=False):
        super().__init__()
        self.logit_scale = torch.nn.Parameter(torch.ones([]) * np.log(1 / 0.07), requires_grad=True)
        var_nfeats = 29 if position_emb else 17
        self.mip_encoder = MIPEncoder(emb_size=emb_dim, var_nfeats=var_nfeats)
        self.sol_encoder = SolutionEncoder(emb_num, emb_dim, n_heads, n_layers)
        self.padding_len = padding_len

    def get_features(self, mip, x=None):
        """-> (mip_features (G,L,128), x_features | None, key_padding_mask (G,L) bool), cmsp.py:35-42."""
        with torch.no_grad():
            mip_features, key_padding_mask = self.encode_mip(mip, mip.n_int_vars)
            x_features = None
            if x is not None:
                x_features, key_padding_mask = self.encode_solution(x, mip.n_int_vars)
        return mip_features, x_features, key_padding_mask

    def encode_solution(self, x, n_int_vars):
        """Pad each instance's 0/1 solution with token 2 to padding_len, then SolutionEncoder (cmsp.py:65-69, utils.get_padding)."""
        sizes = [int(n_int_vars)] if isinstance(n_int_vars, int) else [int(v) for v in n_int_vars.tolist()]
        if sum(sizes) != x.numel():
            raise ValueError("n_int_vars does not add up to len(x)")
        L = self.padding_len
        tok = torch.full((len(sizes), L), 2, dtype=torch.int64, device=x.device)
        mask = torch.ones((len(sizes), L), dtype=torch.bool, device=x.device)
        off = 0
        for g, n in enumerate(sizes):
            if n > L:
                raise ValueError("instance has %d integer variables but padding_len is %d" % (n, L))
            tok[g, :n] = x[off:off + n].to(torch.int64)
            mask[g, :n] = False
            off += n
        return self.sol_encoder(tok, mask), mask

    def encode_mip(self, mip, n_int_vars):
        """GNN -> gather int_indices -> zero pad to padding_len (cmsp.py:53-63, utils.get_padding)."""
        feats = self.mip_encoder(mip.constraint_features, mip.edge_index, mip.edge_attr, mip.variable_features)
        idx = mip.int_indices.to(torch.int64)
        sizes = [int(n_int_vars)] if isinstance(n_int_vars, int) else [int(v) for v in n_int_vars.tolist()]
        if sum(sizes) != idx.numel():
            raise ValueError("n_int_vars does not add up to len(int_indices)")
        L = self.padding_len
        outs, masks, off = [], [], 0
        for n in sizes:
            if n > L:
                raise ValueError("instance has %d integer variables but padding_len is %d" % (n, L))
            outs.append(ops.gather_pad(feats, idx[off:off + n], L).unsqueeze(0))
            m = torch.zeros((1, L), dtype=torch.bool, device=feats.device)
            m[:, n:] = True
            masks.append(m)
            off += n
        return torch.cat(outs, dim=0), torch.cat(masks, dim=0)

    def forward(self, mip, x):
        raise NotImplementedError("contrastive training forward (cmsp.py:65-85) is out of scope of the sampling path")
