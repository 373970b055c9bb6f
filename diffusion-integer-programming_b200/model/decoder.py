"""MipConditionalDecorder (/root/reference/model/decoder.py:10-57) on the sm_100a engine."""
from collections import OrderedDict

import torch

from dip_b200 import ops
from dip_b200.engine import default_engine
from model.modules import ResidualAttentionBlock


import itertools

_module_uid = itertools.count(1)      # identity of a module object that survives address reuse (unlike id())


def weights_signature(module):
    """Changes whenever a parameter tensor is replaced or modified in place (load_state_dict, .to())."""
    return tuple((k, v.data_ptr(), v._version) for k, v in module.state_dict().items())


def tensor_key(t):
    """Identity of a tensor's current content: address, in-place version counter, shape, device."""
    return None if t is None else (t.data_ptr(), t._version, tuple(t.shape), str(t.device), t.dtype)


_same_rows_cache = {}


def same_rows(t):
    """True when every batch row of ``t`` equals row 0 (one instance repeated sample_num times, sample.py:151).  The comparison
    (a full elementwise pass + a host sync) is remembered per tensor identity, so a sampling loop pays for it once."""
    if t.shape[0] == 1:
        return True
    key = tensor_key(t)
    hit = _same_rows_cache.get(key)
    if hit is None:
        hit = (bool((t == t[:1]).all()), t)       # the entry keeps `t` alive so its address cannot be recycled under the key
        if len(_same_rows_cache) >= 8:
            _same_rows_cache.clear()
        _same_rows_cache[key] = hit
    return hit[0]


_cond_cache = OrderedDict()


def use_conditions(eng, mip_row, kpm_row):
    """Make (mip_row, kpm_row) the engine's current constraint-free instance; the prepared form is cached per tensor identity and
    nothing is done at all when the engine already holds it."""
    key = (id(eng), tensor_key(mip_row), tensor_key(kpm_row))
    if getattr(eng, "_current_key", None) == key:
        return
    prep = _cond_cache.get(key)
    if prep is None:
        prep = (eng.prepare_conditions(mip_row, kpm_row), (mip_row, kpm_row))     # keyed tensors kept alive, see same_rows
        _cond_cache[key] = prep
        while len(_cond_cache) > 8:
            _cond_cache.popitem(last=False)
    eng.use_instance(prep[0])
    eng._current_key = key


class MipConditionalDecorder(torch.nn.Module):
    """latent + instance embedding -> per-variable Bernoulli probabilities.

    forward(mip_features (B,L,128), x_features (B,L,128), key_padding_mask (B,L) bool) ->
    (sols, selected): flat ``masked_select`` vectors over the unpadded positions (decoder.py:50-56).
    """

    def __init__(self, attn_dim: int, n_heads: int, n_layers: int, attn_mask: torch.Tensor = None, use_select_net=False):
        super().__init__()
        self.attn_dim = attn_dim
        self.resblocks = torch.nn.ModuleList([ResidualAttentionBlock(attn_dim, n_heads, attn_mask) for _ in range(n_layers)])
        self.linear_proj = torch.nn.Sequential(OrderedDict([("linear_projection", torch.nn.Linear(attn_dim, 1))]))
        self.use_select_net = use_select_net
        if self.use_select_net:
            self.select_net = torch.nn.Sequential(torch.nn.Linear(attn_dim, attn_dim), torch.nn.ReLU(), torch.nn.Linear(attn_dim, 1))
        self._uid = next(_module_uid)

    def sync_engine(self, engine):
        """Upload the weights unless the ENGINE already holds exactly these (ownership is recorded on the engine, so two decoders
        sharing it re-upload when they alternate)."""
        for blk in self.resblocks:
            blk.check_supported()
        sig = (self._uid, weights_signature(self))
        if engine.dec_sig != sig:
            engine.set_decoder({k: v.detach() for k, v in self.state_dict().items()})
            engine.dec_sig = sig

    def probabilities(self, mip_features, x_features, key_padding_mask, want_hidden=False):
        """(B,L) probabilities with exact zeros on padding (+ the last-L hidden state)."""
        if key_padding_mask is None:
            key_padding_mask = torch.zeros(mip_features.shape[:2], dtype=torch.bool, device=mip_features.device)
        eng = default_engine(x_features.device)
        self.sync_engine(eng)
        x = x_features.detach().to(torch.float32).contiguous()
        if same_rows(mip_features) and same_rows(key_padding_mask):
            use_conditions(eng, mip_features[0], key_padding_mask[0])
            return eng.decode(x, want_hidden=want_hidden)
        ps, ys = [], []
        for i in range(x.shape[0]):     # distinct instances in one batch: one wave each
            use_conditions(eng, mip_features[i], key_padding_mask[i])
            r = eng.decode(x[i:i + 1], want_hidden=want_hidden)
            ps.append(r[0] if want_hidden else r)
            if want_hidden:
                ys.append(r[1])
        return (torch.cat(ps), torch.cat(ys)) if want_hidden else torch.cat(ps)

    def forward(self, mip_features: torch.Tensor, x_features: torch.Tensor, key_padding_mask: torch.Tensor = None):
        if x_features.requires_grad:
            raise NotImplementedError("autograd through the decoder is replaced by the engine's input-gradient kernels "
                                      "(Engine.guidance_grad); torch autograd is not supported on this path")
        valid = None if key_padding_mask is None else ~key_padding_mask
        if not self.use_select_net:
            p = self.probabilities(mip_features, x_features, key_padding_mask)
            return (p.reshape(-1) if valid is None else torch.masked_select(p, valid)), None
        p, y = self.probabilities(mip_features, x_features, key_padding_mask, want_hidden=True)
        B, L, Dm = y.shape
        w = {k: v.detach().contiguous() for k, v in self.select_net.state_dict().items()}
        h = ops.gemm_nt(y.view(B * L, Dm), w["0.weight"], w["0.bias"], act="relu")
        kpm = key_padding_mask if key_padding_mask is not None else torch.zeros((B, L), dtype=torch.bool, device=y.device)
        sel = ops.decoder_head_fwd(h.view(B, L, Dm), B, L, L, w["2.weight"].view(-1), w["2.bias"], kpm)       # one mask row per sample
        if valid is None:
            return p.reshape(-1), sel.reshape(-1)
        return torch.masked_select(p, valid), torch.masked_select(sel, valid)


class SelectiveNetLoss:
    """Training loss of the selective decoder (reference: decoder.py:60-75): out of scope."""

    def __init__(self, lamb=32, coverage=0.50):
        raise NotImplementedError("SelectiveNetLoss belongs to training; out of scope of the sampling path")
