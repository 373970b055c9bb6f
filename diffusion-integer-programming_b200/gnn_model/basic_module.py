"""The reference keeps a second copy of its GNN classes here (/root/reference/gnn_model/basic_module.py:127-235 is identical to
model/modules.py:104-212); this mirror re-exports the one set of parameter containers."""
from model.modules import (BipartiteGraphConvolution, GCNMLPModule, PreNormException, PreNormLayer,  # noqa: F401
                           SelectiveNet)


class BaseModel:            # the reference's BaseModel only adds pre-training helpers (training: out of scope)
    pass
