"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Stand-ins for the two third-party packages the reference imports at module top but that are
absent from this image:

* ``torch_sparse.spmm`` (rusty1s/pytorch_sparse, version unpinned by the reference; call sites
  /root/reference/model/diffusion.py:624 and /root/reference/sample_partialsol.py:167).
  Published semantics restated here: ``out = matrix.index_select(-2, col) * value[:, None]``
  followed by a scatter-add along dim -2 into ``m`` rows.
* ``torch_geometric.nn.MessagePassing('add')`` (pyg-team/pytorch_geometric, unpinned; call site
  /root/reference/model/modules.py:173-212).  Semantics: flow source->target,
  ``x_j = feats[0][edge_index[0]]``, ``x_i = feats[1][edge_index[1]]``, messages summed at
  ``edge_index[1]`` into ``size[1]`` rows.
* ``torch_geometric.nn.aggr.{Mean,Sum}Aggregation``, ``torch_geometric.data.{Data,Dataset}``:
  constructed by the reference (model/encoder.py:16, model/decoder.py:63-64) but never called on the
  sampling path -- dummies.

``install()`` registers them in ``sys.modules`` so the UNMODIFIED reference files import.
"""
import sys
import types

import torch


def spmm(index, value, m, n, matrix):
    """torch_sparse.spmm restatement (supports a leading batch dim on ``matrix``)."""
    row, col = index[0], index[1]
    matrix = matrix if matrix.dim() > 1 else matrix.unsqueeze(-1)
    out = matrix.index_select(-2, col)
    out = out * value.unsqueeze(-1)
    shape = list(out.shape)
    shape[-2] = m
    res = torch.zeros(shape, dtype=out.dtype, device=out.device)
    res.index_add_(out.dim() - 2, row, out)
    return res


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", **kwargs):
        super().__init__()
        assert aggr == "add"

    def propagate(self, edge_index, size=None, **kwargs):
        node_features = kwargs["node_features"]
        edge_features = kwargs["edge_features"]
        left, right = node_features
        x_j = left[edge_index[0]]
        x_i = right[edge_index[1]]
        msg = self.message(x_i, x_j, edge_features)
        out = torch.zeros((size[1], msg.shape[-1]), dtype=msg.dtype, device=msg.device)
        out.index_add_(0, edge_index[1], msg)
        return out


class _DummyAggr(torch.nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - never on the sampling path
        raise RuntimeError("aggregation dummy called; not on the sampling path")


class _Data:
    def __init__(self, *a, **k):
        pass


def install():
    if "torch_sparse" not in sys.modules:
        ts = types.ModuleType("torch_sparse")
        ts.spmm = spmm
        sys.modules["torch_sparse"] = ts
    if "torch_geometric" not in sys.modules:
        tg = types.ModuleType("torch_geometric")
        nn = types.ModuleType("torch_geometric.nn")
        aggr = types.ModuleType("torch_geometric.nn.aggr")
        data = types.ModuleType("torch_geometric.data")
        loader = types.ModuleType("torch_geometric.loader")
        nn.MessagePassing = MessagePassing
        aggr.MeanAggregation = _DummyAggr
        aggr.SumAggregation = _DummyAggr
        nn.aggr = aggr
        data.Data = _Data
        data.Dataset = _Data
        loader.DataLoader = _Data
        tg.nn = nn
        tg.data = data
        tg.loader = loader
        sys.modules["torch_geometric"] = tg
        sys.modules["torch_geometric.nn"] = nn
        sys.modules["torch_geometric.nn.aggr"] = aggr
        sys.modules["torch_geometric.data"] = data
        sys.modules["torch_geometric.loader"] = loader
