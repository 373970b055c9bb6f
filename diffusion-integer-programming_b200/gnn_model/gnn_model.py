"""NeuralDiving head on the shared K3 encoder (/root/reference/gnn_model/gnn_model.py:61-86): the competing baseline of the
reference's plots (contrast_plot.py), run on the same sm_100a kernels as the CMSP encoder -- forward only."""
import torch

from dip_b200 import ops
from model.modules import GNNModel, SelectiveNet  # noqa: F401  (GNNModel re-exported: gnn_model.py:5-59 duplicates modules.py:214-268)


class NeuralDiving(torch.nn.Module):
    """variable embeddings -> (Bernoulli probability per variable, SelectiveNet coverage score per variable).
    Same constructor, module tree and state_dict keys as the reference."""

    def __init__(self, emb_size=64, cons_nfeats=5, edge_nfeats=1, var_nfeats=17, gcn_mlp_layer_num=5):
        super().__init__()
        self.gnn_model = GNNModel(emb_size, cons_nfeats, edge_nfeats, var_nfeats, gcn_mlp_layer_num=gcn_mlp_layer_num)
        self.output_module = torch.nn.Sequential(torch.nn.Linear(emb_size, emb_size), torch.nn.ReLU(), torch.nn.Linear(emb_size, 1),
                                                 torch.nn.Sigmoid())
        self.select_net = SelectiveNet(emb_size)

    @torch.no_grad()
    def forward(self, constraint_features, edge_indices, edge_features, variable_features):
        v = self.gnn_model(constraint_features, edge_indices, edge_features, variable_features)        # (N, 128), K3 kernels
        N = v.shape[0]
        none = torch.zeros(N, dtype=torch.bool, device=v.device)
        w = {k: t.detach().contiguous() for k, t in self.state_dict().items()}
        h = ops.gemm_nt(v, w["output_module.0.weight"], w["output_module.0.bias"], act="relu")
        out = ops.decoder_head_fwd(h.view(1, N, -1), 1, N, N, w["output_module.2.weight"].view(-1), w["output_module.2.bias"], none)
        hs = ops.gemm_nt(v, w["select_net.select_module.0.weight"], w["select_net.select_module.0.bias"], act="relu")
        zero = torch.zeros(1, dtype=torch.float32, device=v.device)
        sel = ops.decoder_head_fwd(hs.view(1, N, -1), 1, N, N, w["select_net.select_module.2.weight"].view(-1), zero, none)
        return out.view(N), sel.view(N)
