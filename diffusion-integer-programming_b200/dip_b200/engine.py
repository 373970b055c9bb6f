"""Python handle on the C engine (include/dip_b200.h, "engine" section).

PyTorch is plumbing here: it owns device memory (weights, latents, workspace) and the CUDA stream;
every computation of the sampling path is a kernel of libdip_b200.so.  Tensors handed to the
engine are borrowed by raw pointer, so the handle keeps references to them.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr

D = 128

BLOCK_KEYS = [
    ("ln1_w", "ln_1.weight"), ("ln1_b", "ln_1.bias"), ("in_proj_w", "attn.in_proj_weight"), ("in_proj_b", "attn.in_proj_bias"),
    ("out_proj_w", "attn.out_proj.weight"), ("out_proj_b", "attn.out_proj.bias"), ("ln2_w", "ln_2.weight"), ("ln2_b", "ln_2.bias"),
    ("fc_w", "mlp.c_fc.weight"), ("fc_b", "mlp.c_fc.bias"), ("proj_w", "mlp.c_proj.weight"), ("proj_b", "mlp.c_proj.bias"),
]


def _stream(device=None):
    return torch.cuda.current_stream(device).cuda_stream


def _dev(t, dtype=torch.float32, device=None):
    """contiguous tensor of `dtype` on the CUDA device (numpy / torch accepted)."""
    if isinstance(t, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(t))
    t = t.detach()
    if device is None:
        device = t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())
    return t.to(device=device, dtype=dtype).contiguous()


def _on_device(fn):
    """Run an Engine method with the engine's device current (kernels launch on the current device's stream)."""
    import functools

    @functools.wraps(fn)
    def wrapper(self, *a, **k):
        if torch.cuda.current_device() == self.device.index:
            return fn(self, *a, **k)
        with torch.cuda.device(self.device):
            return fn(self, *a, **k)
    return wrapper


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.DipError("diffusion-integer-programming_b200 needs a CUDA device (sm_100a); there is no CPU path")


def coo_to_csr_csc(rows, cols, vals, M, n):
    """Host CSR + CSC of the coalesced matrix via the library's own builder (numpy in / out)."""
    lib = _lib.load()
    rows = np.ascontiguousarray(rows, dtype=np.int64)
    cols = np.ascontiguousarray(cols, dtype=np.int64)
    vals = np.ascontiguousarray(vals, dtype=np.float32)
    nnz = len(rows)
    out = {
        "csr_rowptr": np.zeros(M + 1, np.int32), "csr_col": np.zeros(max(nnz, 1), np.int32), "csr_val": np.zeros(max(nnz, 1), np.float32),
        "csc_colptr": np.zeros(n + 1, np.int32), "csc_row": np.zeros(max(nnz, 1), np.int32), "csc_val": np.zeros(max(nnz, 1), np.float32),
    }
    nnz_out = C.c_int64(0)
    check(lib.dip_coo_to_csr_csc_host(nnz, ptr(rows), ptr(cols), ptr(vals), M, n, ptr(out["csr_rowptr"]), ptr(out["csr_col"]),
                                      ptr(out["csr_val"]), ptr(out["csc_colptr"]), ptr(out["csc_row"]), ptr(out["csc_val"]),
                                      C.byref(nnz_out)))
    m = nnz_out.value
    for k in ("csr_col", "csr_val", "csc_row", "csc_val"):
        out[k] = out[k][:max(m, 1)]
    out["nnz"] = m
    return out


class Engine:
    """One engine per process / GPU (the reference's samplers are not re-entrant either)."""

    def __init__(self, device=None):
        require_cuda()
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        with torch.cuda.device(self.device):
            h = C.c_void_p()
            check(self.lib.dip_engine_create(C.byref(h)))
        self.h = h
        self._keep = {}
        self._ws = None
        self.inst = None
        self.n_den = self.n_dec = 0
        self.precision = "tc"
        # which module's weights the engine holds (set by the model.* mirrors' sync_engine): ownership lives HERE, so two model
        # objects sharing the engine re-upload when they alternate
        self.den_sig = None
        self.dec_sig = None
        self._current_key = None      # identity of the current instance as the model.* mirrors key it (reset by use_instance)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.dip_engine_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def set_precision(self, mode):
        """'tc' (default: tcgen05 tensor cores, bf16x2-split operands) or 'fp32' (exact SIMT kernels)."""
        code = {"tc": 1, "fp32": 0, 1: 1, 0: 0}[mode]
        check(self.lib.dip_engine_set_precision(self.h, code))
        self.precision = "tc" if code == 1 else "fp32"

    # ---- weights -------------------------------------------------------------------------------
    def _blocks(self, sd, prefix, n_layers, keep):
        arr = (_lib.BlockWeights * max(n_layers, 1))()
        for i in range(n_layers):
            for field, key in BLOCK_KEYS:
                t = _dev(sd["%s%d.%s" % (prefix, i, key)], device=self.device)
                keep.append(t)
                setattr(arr[i], field, t.data_ptr())
        return arr

    @staticmethod
    def count_layers(sd, prefix):
        i = 0
        while ("%s%d.ln_1.weight" % (prefix, i)) in sd:
            i += 1
        return i

    def set_denoiser(self, sd, prefix="predict_model."):
        """sd: state_dict-like of DDPMTrainer (keys ``predict_model.*``)."""
        keep = []
        n = self.count_layers(sd, prefix + "resblocks.")
        arr = self._blocks(sd, prefix + "resblocks.", n, keep)
        w1, b1, w2, b2 = (_dev(sd[prefix + k], device=self.device) for k in
                          ("mlp_xt.linear_1.weight", "mlp_xt.linear_1.bias", "mlp_xt.linear_2.weight", "mlp_xt.linear_2.bias"))
        if tuple(w1.shape) != (D, 2 * D) or tuple(w2.shape) != (D, D):
            raise _lib.DipError("denoiser width must be 128 (got linear_1 %s)" % (tuple(w1.shape),))
        with torch.cuda.device(self.device):
            check(self.lib.dip_engine_set_denoiser(self.h, ptr(w1), ptr(b1), ptr(w2), ptr(b2), arr, n, _stream()))
            torch.cuda.current_stream().synchronize()   # weights were copied; sources may go away
        self.n_den = n
        self.den_sig = None

    def set_decoder(self, sd):
        keep = []
        n = self.count_layers(sd, "resblocks.")
        arr = self._blocks(sd, "resblocks.", n, keep)
        pw = _dev(sd["linear_proj.linear_projection.weight"], device=self.device)
        pb = _dev(sd["linear_proj.linear_projection.bias"], device=self.device)
        if tuple(pw.shape) != (1, D):
            raise _lib.DipError("decoder width must be 128")
        with torch.cuda.device(self.device):
            check(self.lib.dip_engine_set_decoder(self.h, arr, n, ptr(pw), ptr(pb), _stream()))
            torch.cuda.current_stream().synchronize()
        self.n_dec = n
        self.dec_sig = None

    # ---- instance ------------------------------------------------------------------------------
    def prepare_instance(self, mip, kpm, rows, cols, vals, M, b, c, a_int=None, b_int=None, c_int=None):
        """Build the device-side form of one instance ONCE (host CSR/CSC build by the library's deterministic coalescing sort, uploads,
        the dip_instance struct): mip (L,128) conditions of ONE instance; kpm (L,) bool; A as COO (rows, cols, vals) of shape
        (M, n) with n = number of unmasked positions; b (M,), c (n,).  Optional exact-integer twin a_int (aligned with
        rows/cols/vals, duplicates not allowed then), b_int, c_int.  Returns a handle for use_instance()."""
        dev = self.device
        mip = _dev(mip, device=dev)
        L = mip.shape[0]
        kpm_t = _dev(kpm, torch.bool, device=dev).view(-1)
        if kpm_t.numel() != L:
            raise _lib.DipError("key_padding_mask length %d != padding_len %d" % (kpm_t.numel(), L))
        pos = torch.nonzero(~kpm_t).view(-1).to(torch.int32).contiguous()
        n = int(pos.numel())
        c = _dev(c, device=dev).view(-1)
        b = _dev(b, device=dev).view(-1)
        if c.numel() != n:
            # the reference's `pred_x.squeeze() @ c` raises for the same mismatch
            raise _lib.DipError("objective vector has %d entries but %d variables are unmasked" % (c.numel(), n))
        if b.numel() != M:
            raise _lib.DipError("rhs has %d entries, A has %d rows" % (b.numel(), M))
        to_np = lambda t: t.detach().cpu().numpy() if torch.is_tensor(t) else np.asarray(t)
        rows_np, cols_np, vals_np = to_np(rows), to_np(cols), to_np(vals)
        csr = coo_to_csr_csc(rows_np, cols_np, vals_np, M, n)
        keep = {"mip": mip, "kpm": kpm_t.to(torch.uint8).contiguous(), "pos": pos, "b": b, "c": c}
        for k in ("csr_rowptr", "csr_col", "csr_val", "csc_colptr", "csc_row", "csc_val"):
            keep[k] = torch.from_numpy(csr[k]).to(dev)
        inst = _lib.Instance()
        inst.L, inst.n, inst.M, inst.nnz = L, n, int(M), int(csr["nnz"])
        if a_int is not None:
            if csr["nnz"] != len(rows_np):
                raise _lib.DipError("integer twin needs a duplicate-free COO")
            ci = coo_to_csr_csc(rows_np, cols_np, np.asarray(to_np(a_int), dtype=np.float32), M, n)
            keep["csr_val_int"] = torch.from_numpy(np.rint(ci["csr_val"]).astype(np.int32)).to(dev)
            keep["b_int"] = _dev(np.asarray(to_np(b_int), dtype=np.int32), torch.int32, device=dev)
            keep["c_int"] = _dev(np.asarray(to_np(c_int), dtype=np.int32), torch.int32, device=dev)
        for f, k in (("mip", "mip"), ("kpm", "kpm"), ("pos_of_col", "pos"), ("csr_rowptr", "csr_rowptr"), ("csr_col", "csr_col"),
                     ("csr_val", "csr_val"), ("csc_colptr", "csc_colptr"), ("csc_row", "csc_row"), ("csc_val", "csc_val"),
                     ("b", "b"), ("c", "c"), ("csr_val_int", "csr_val_int"), ("b_int", "b_int"), ("c_int", "c_int")):
            setattr(inst, f, keep[k].data_ptr() if k in keep else None)
        return {"struct": inst, "keep": keep,
                "info": {"L": L, "n": n, "M": int(M), "nnz": int(csr["nnz"]), "has_int": a_int is not None}}

    def use_instance(self, prepared):
        """Make a prepared instance the engine's current one (no host work beyond one library call)."""
        with torch.cuda.device(self.device):
            check(self.lib.dip_engine_set_instance(self.h, C.byref(prepared["struct"]), _stream(self.device)))
        self._current_key = None
        self._keep["inst"] = prepared            # the library borrows the device arrays: keep them alive
        self.inst = prepared["info"]
        return self.inst

    def use_instances(self, prepared_list, inst_of=None):
        """Several prepared instances of one family resident at once + (optionally) the wave map: inst_of[s] = index into
        `prepared_list` of the instance sample s of the following waves belongs to (sample_partialsol.py draws 15 samples per
        instance: a wave then mixes several instances instead of leaving the GPU two thirds idle).  Per-sample outputs of width n
        become n_max wide, zero padded."""
        if len(prepared_list) == 0:
            raise _lib.DipError("use_instances: empty list")
        arr = (_lib.Instance * len(prepared_list))(*[p["struct"] for p in prepared_list])
        with torch.cuda.device(self.device):
            check(self.lib.dip_engine_set_instances(self.h, arr, len(prepared_list), _stream(self.device)))
        self._current_key = None
        self._keep["inst"] = list(prepared_list)
        infos = [p["info"] for p in prepared_list]
        self.inst = {"L": infos[0]["L"], "n": max(i["n"] for i in infos), "M": max(i["M"] for i in infos),
                     "nnz": sum(i["nnz"] for i in infos), "has_int": all(i["has_int"] for i in infos), "n_inst": len(infos),
                     "n_of": [i["n"] for i in infos]}
        if inst_of is not None:
            self.set_wave_map(inst_of)
        return self.inst

    def set_wave_map(self, inst_of):
        """inst_of: sequence / array of instance indices, one per sample of the waves that follow (None clears the map)."""
        with torch.cuda.device(self.device):
            if inst_of is None:
                check(self.lib.dip_engine_set_wave_map(self.h, None, 0, _stream(self.device)))
                return
            m = np.ascontiguousarray(np.asarray(inst_of, dtype=np.int32))
            check(self.lib.dip_engine_set_wave_map(self.h, ptr(m), int(m.size), _stream(self.device)))
            torch.cuda.current_stream().synchronize()           # the host array is read by an async copy

    def set_instance(self, mip, kpm, rows, cols, vals, M, b, c, a_int=None, b_int=None, c_int=None):
        """prepare_instance + use_instance."""
        return self.use_instance(self.prepare_instance(mip, kpm, rows, cols, vals, M, b, c, a_int=a_int, b_int=b_int, c_int=c_int))

    def prepare_conditions(self, mip, kpm):
        """Instance without constraints (unguided sampling / plain decoding): an empty 1-row system."""
        kpm = kpm.view(-1)
        n = int((~kpm.to(torch.bool)).sum())
        z = np.zeros(0, dtype=np.int64)
        return self.prepare_instance(mip, kpm, z, z, np.zeros(0, dtype=np.float32), 1, np.zeros(1, dtype=np.float32),
                                     np.zeros(n, dtype=np.float32))

    def set_conditions(self, mip, kpm):
        return self.use_instance(self.prepare_conditions(mip, kpm))

    # ---- workspace -----------------------------------------------------------------------------
    @_on_device
    def workspace(self, B):
        need = self.lib.dip_engine_workspace_bytes(self.h, int(B))
        if need == 0:
            raise _lib.DipError("set_instance must be called before sampling")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need + 256, dtype=torch.uint8, device=self.device)
        base = self._ws.data_ptr()
        off = (-base) % 256
        return base + off, self._ws.numel() - off

    def reset_cache(self):
        """Forget the wave constants kept in the workspace (needed only if something else wrote into the workspace tensor)."""
        check(self.lib.dip_engine_reset_cache(self.h))

    def _x(self, x):
        if not (torch.is_tensor(x) and x.is_cuda and x.dtype == torch.float32 and x.is_contiguous()):
            raise _lib.DipError("latents must be a contiguous fp32 CUDA tensor")
        if x.device != self.device:
            raise _lib.DipError("latents live on %s but this engine was created on %s" % (x.device, self.device))
        if x.dim() != 3 or x.shape[1] != self.inst["L"] or x.shape[2] != D:
            raise _lib.DipError("latents must be (B, %d, 128), got %s" % (self.inst["L"], tuple(x.shape)))
        return x

    # ---- networks ------------------------------------------------------------------------------
    @_on_device
    def denoise(self, x, t):
        x = self._x(x)
        out = torch.empty_like(x)
        ws, nb = self.workspace(x.shape[0])
        check(self.lib.dip_engine_denoise(self.h, ptr(out), ptr(x), x.shape[0], float(t), ws, nb, _stream()))
        return out

    @_on_device
    def decode(self, x, want_hidden=False):
        x = self._x(x)
        B = x.shape[0]
        p = torch.empty((B, self.inst["L"]), dtype=torch.float32, device=x.device)
        y = torch.empty_like(x) if want_hidden else None
        ws, nb = self.workspace(B)
        check(self.lib.dip_engine_decode(self.h, ptr(p), ptr(y), ptr(x), B, ws, nb, _stream()))
        return (p, y) if want_hidden else p

    @_on_device
    def guidance_grad(self, x, lam):
        x = self._x(x)
        B = x.shape[0]
        grad = torch.empty_like(x)
        p = torch.empty((B, self.inst["L"]), dtype=torch.float32, device=x.device)
        viol = torch.empty(B, dtype=torch.float32, device=x.device)
        obj = torch.empty(B, dtype=torch.float32, device=x.device)
        ws, nb = self.workspace(B)
        check(self.lib.dip_engine_guidance_grad(self.h, ptr(grad), ptr(p), ptr(viol), ptr(obj), ptr(x), B, float(lam), ws, nb, _stream()))
        return grad, p, viol, obj

    # ---- sampler steps -------------------------------------------------------------------------
    @staticmethod
    def ddim_coeffs(t, sqrt_at, s1m, c_dir, sqrt_aprev, sigma):
        return _lib.DdimCoeffs(float(t), float(sqrt_at), float(s1m), float(c_dir), float(sqrt_aprev), float(sigma))

    @_on_device
    def guided_ddim_step(self, x, co, gs, lam, eps_param=False, noise=None, x0_out=None):
        x = self._x(x)
        ws, nb = self.workspace(x.shape[0])
        check(self.lib.dip_engine_guided_ddim_step(self.h, ptr(x), ptr(x0_out), x.shape[0], C.byref(co), float(gs), float(lam),
                                                   int(eps_param), ptr(noise), ws, nb, _stream()))
        return x

    @_on_device
    def ddim_step(self, x, co, eps_param=False, noise=None, x0_out=None):
        x = self._x(x)
        ws, nb = self.workspace(x.shape[0])
        check(self.lib.dip_engine_ddim_step(self.h, ptr(x), ptr(x0_out), x.shape[0], C.byref(co), int(eps_param), ptr(noise), ws, nb, _stream()))
        return x

    @_on_device
    def guided_ddpm_step(self, x, co, gs, lam, eps_param=False, noise=None):
        x = self._x(x)
        ws, nb = self.workspace(x.shape[0])
        check(self.lib.dip_engine_guided_ddpm_step(self.h, ptr(x), x.shape[0], C.byref(co), float(gs), float(lam), int(eps_param),
                                                   ptr(noise), ws, nb, _stream()))
        return x

    @_on_device
    def ddpm_step(self, x, co, eps_param=False, noise=None):
        x = self._x(x)
        ws, nb = self.workspace(x.shape[0])
        check(self.lib.dip_engine_ddpm_step(self.h, ptr(x), x.shape[0], C.byref(co), int(eps_param), ptr(noise), ws, nb, _stream()))
        return x

    @_on_device
    def final_decode(self, x, want_p=True):
        """Decode + round + feasibility (sample.py:162-174).  Returns dict of device tensors."""
        x = self._x(x)
        B, n, L = x.shape[0], self.inst["n"], self.inst["L"]
        dev = x.device
        out = {
            "p_full": torch.empty((B, L), dtype=torch.float32, device=dev) if want_p else None,
            "xbits": torch.empty((B, n), dtype=torch.uint8, device=dev),
            "penalty": torch.empty(B, dtype=torch.float32, device=dev),
            "viol_int": torch.empty(B, dtype=torch.int64, device=dev),
            "obj_int": torch.empty(B, dtype=torch.int64, device=dev),
        }
        ws, nb = self.workspace(B)
        check(self.lib.dip_engine_final_decode(self.h, ptr(x), B, ptr(out["p_full"]), ptr(out["xbits"]), ptr(out["penalty"]),
                                               ptr(out["viol_int"]), ptr(out["obj_int"]), ws, nb, _stream()))
        return out

    @_on_device
    def sample_guided_ddim(self, x, steps, gs, lam, eps_param=False):
        """Whole S-step loop + final decode in ONE library call (in place on x)."""
        x = self._x(x)
        B, n = x.shape[0], self.inst["n"]
        dev = x.device
        arr = (_lib.DdimCoeffs * len(steps))(*steps)
        out = {
            "xbits": torch.empty((B, n), dtype=torch.uint8, device=dev),
            "penalty": torch.empty(B, dtype=torch.float32, device=dev),
            "viol_int": torch.empty(B, dtype=torch.int64, device=dev),
            "obj_int": torch.empty(B, dtype=torch.int64, device=dev),
        }
        ws, nb = self.workspace(B)
        check(self.lib.dip_engine_sample_guided_ddim(self.h, ptr(x), B, arr, len(steps), float(gs), float(lam), int(eps_param),
                                                     ptr(out["xbits"]), ptr(out["penalty"]), ptr(out["viol_int"]), ptr(out["obj_int"]),
                                                     ws, nb, _stream()))
        out["x"] = x
        return out


_default = {}


def default_engine(device=None):
    """Per-device engine shared by the drop-in ``model.*`` classes (device: the tensors' CUDA device; default: the current one)."""
    require_cuda()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    if dev not in _default:
        _default[dev] = Engine(dev)
    return _default[dev]


def randn(shape, seed, subseq, device=None):
    """Counter-based N(0,1) tensor from the library's Philox generator."""
    require_cuda()
    lib = _lib.load()
    out = torch.empty(shape, dtype=torch.float32, device=device or torch.device("cuda", torch.cuda.current_device()))
    check(lib.dip_randn(ptr(out), out.numel(), int(seed), int(subseq), _stream()))
    return out
