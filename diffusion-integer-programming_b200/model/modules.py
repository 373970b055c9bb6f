"""Parameter containers with the reference's module tree (/root/reference/model/modules.py), so
that the reference's checkpoints load unchanged.  Forward passes on the sampling path are executed
by libdip_b200.so; nothing here runs torch math on the hot path.
"""
from collections import OrderedDict

import torch

from dip_b200 import ops


class PreNormException(Exception):
    pass


class PreNormLayer(torch.nn.Module):
    """(x + shift) * scale with frozen buffers (reference: modules.py:15-37).  The pre-training
    statistics pass (start_updates / update_stats / stop_updates) belongs to training: out of scope."""

    def __init__(self, n_units, shift=True, scale=True, name=None):
        super().__init__()
        assert shift or scale
        self.register_buffer("shift", torch.zeros(n_units) if shift else None)
        self.register_buffer("scale", torch.ones(n_units) if scale else None)
        self.n_units = n_units
        self.waiting_updates = False
        self.received_updates = False


class QuickGELU(torch.nn.Module):
    """x * sigmoid(1.702 x) (reference: modules.py:270-272); fused into the GEMM epilogues."""

    def forward(self, x):
        raise NotImplementedError("QuickGELU is fused into the CUDA GEMM epilogue; call the owning block")


class ResidualAttentionBlock(torch.nn.Module):
    """Pre-LN attention + D->D->D QuickGELU MLP, both residual (reference: modules.py:274-296)."""

    def __init__(self, d_model: int, n_head: int, attn_mask: torch.Tensor = None):
        super().__init__()
        self.attn = torch.nn.MultiheadAttention(d_model, n_head, batch_first=True)
        self.ln_1 = torch.nn.LayerNorm(d_model)
        self.mlp = torch.nn.Sequential(OrderedDict([
            ("c_fc", torch.nn.Linear(d_model, d_model)),
            ("gelu", QuickGELU()),
            ("c_proj", torch.nn.Linear(d_model, d_model)),
        ]))
        self.ln_2 = torch.nn.LayerNorm(d_model)
        self.attn_mask = attn_mask
        self.d_model, self.n_head = d_model, n_head

    def check_supported(self):
        if self.d_model != 128 or self.n_head != 1 or self.attn_mask is not None:
            raise NotImplementedError("the sm_100a kernels are built for d_model=128, n_head=1, attn_mask=None "
                                      "(the configuration of sample.py:21-37)")

    def forward(self, x: torch.Tensor, key_padding_mask: torch.Tensor = None):
        self.check_supported()
        w = {k: v.detach() for k, v in self.state_dict().items()}
        return ops.block_forward(x, w, key_padding_mask)


class SelectiveNet(torch.nn.Module):
    """Parameter container only (reference: modules.py:86-102); unused on the sampling path."""

    def __init__(self, feature_size):
        super().__init__()
        self.select_module = torch.nn.Sequential(
            torch.nn.Linear(feature_size, feature_size), torch.nn.ReLU(),
            torch.nn.Linear(feature_size, 1, bias=False), torch.nn.Sigmoid())


class GCNMLPModule(torch.nn.Module):
    """Z = A f(U) + jump network of Neural Diving (reference: modules.py:104-171): parameters only,
    the forward lives in dip_b200.encoder."""

    def __init__(self, cons_input_size=64, vars_input_size=64, output_size=64):
        super().__init__()
        mlp = lambda i, o: torch.nn.Sequential(torch.nn.Linear(i, o), torch.nn.ReLU(), torch.nn.Linear(o, o), torch.nn.ReLU())
        self.cons_mlp_layer = mlp(cons_input_size, output_size)
        self.vars_mlp_layer = mlp(vars_input_size, output_size)
        self.jump_vars_mlp_layer = torch.nn.Sequential(torch.nn.Linear(output_size * 2, output_size), torch.nn.Sigmoid())
        self.jump_cons_mlp_layer = torch.nn.Sequential(torch.nn.Linear(output_size * 2, output_size), torch.nn.Sigmoid())
        self.vars_layer_norm = torch.nn.LayerNorm(output_size)
        self.cons_layer_norm = torch.nn.LayerNorm(output_size)


class BipartiteGraphConvolution(torch.nn.Module):
    """Half-convolution parameters (reference: modules.py:173-201).  The reference derives from PyG's
    MessagePassing('add'); here the gather / segment-reduce is dip_conv_reduce (dip_b200.encoder)."""

    def __init__(self, emb_size=64):
        super().__init__()
        self.feature_module_left = torch.nn.Sequential(torch.nn.Linear(emb_size, emb_size))
        self.feature_module_edge = torch.nn.Sequential(torch.nn.Linear(1, emb_size, bias=False))
        self.feature_module_right = torch.nn.Sequential(torch.nn.Linear(emb_size, emb_size, bias=False))
        self.feature_module_final = torch.nn.Sequential(PreNormLayer(1, shift=False), torch.nn.ReLU(),
                                                        torch.nn.Linear(emb_size, emb_size))
        self.post_conv_module = torch.nn.Sequential(PreNormLayer(1, shift=False))
        self.output_module = torch.nn.Sequential(torch.nn.Linear(2 * emb_size, emb_size), torch.nn.ReLU(),
                                                 torch.nn.Linear(emb_size, emb_size))


class GNNModel(torch.nn.Module):
    """Bipartite GNN parameters (reference: modules.py:214-247) + forward through the K3 kernels."""

    def __init__(self, emb_size=64, cons_nfeats=5, edge_nfeats=1, var_nfeats=17, gcn_mlp_layer_num=5):
        super().__init__()
        self.gcn_mlp_layer_num = gcn_mlp_layer_num
        self.emb_size = emb_size
        emb = lambda nf: torch.nn.Sequential(PreNormLayer(nf), torch.nn.Linear(nf, emb_size), torch.nn.ReLU(),
                                             torch.nn.Linear(emb_size, emb_size), torch.nn.ReLU())
        self.cons_embedding = emb(cons_nfeats)
        self.edge_embedding = torch.nn.Sequential(PreNormLayer(edge_nfeats))
        self.var_embedding = emb(var_nfeats)
        self.conv_v_to_c = BipartiteGraphConvolution(emb_size)
        self.conv_c_to_v = BipartiteGraphConvolution(emb_size)
        self.gcn_mlp_layer = torch.nn.ModuleList([GCNMLPModule(emb_size, emb_size, emb_size) for _ in range(gcn_mlp_layer_num)])

    def forward(self, constraint_features, edge_indices, edge_attrs, variable_features):
        from dip_b200 import encoder
        if self.emb_size != 128:
            raise NotImplementedError("the sm_100a kernels are built for emb_size=128 (model/cmsp.py:18)")
        w = {k: v.detach() for k, v in self.state_dict().items()}
        return encoder.gnn_forward(w, constraint_features, edge_indices, edge_attrs, variable_features)
