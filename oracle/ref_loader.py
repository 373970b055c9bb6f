"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Imports the UNMODIFIED reference modules from /root/reference (present only in the build
container, never on the GPU box) behind the shims of ``oracle/shims.py`` and hands them back
as a namespace.  The reference's top-level package is called ``model`` -- the same name as
this repo's drop-in mirror -- so the reference modules are imported, captured and then removed
from ``sys.modules`` again; both can live in one process.

Used by ``oracle/gen_golden.py`` (fixture generation) and by the CPU tests that validate
``oracle/restate.py`` against the real reference when it is available.
"""
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference_model.zip")
_unpacked = None


def _unpack():
    """oracle/_ref/reference_model.zip (byte copies staged by oracle/stage_ref.py) -> a temporary directory, once per process."""
    global _unpacked
    if _unpacked is None:
        import atexit
        import shutil
        import tempfile
        import zipfile
        _unpacked = tempfile.mkdtemp(prefix="dip_ref_")
        atexit.register(shutil.rmtree, _unpacked, ignore_errors=True)
        with zipfile.ZipFile(_STAGED) as z:
            z.extractall(_unpacked)
    return _unpacked


def _root():
    """/root/reference in the build container; on the GPU box the staged archive (oracle/_ref/)."""
    env = os.environ.get("DIP_REFERENCE_ROOT")
    for cand in ([env] if env else []) + ["/root/reference"]:
        if env == "staged":           # force the staged archive (what the GPU box sees)
            break
        if os.path.isfile(os.path.join(cand, "model", "diffusion.py")):
            return cand
    if os.path.isfile(_STAGED):
        return _unpack()
    return "/root/reference"


REFERENCE_ROOT = _root()


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model", "diffusion.py"))


def kind():
    """'reference' when the unmodified files are importable (in place or staged), else None."""
    return "reference" if available() else None


_cache = None


def load():
    """Return a namespace with the reference's classes (DDIMSampler, DDPMSampler, DDPMTrainer,
    MipConditionalDecorder, CMSP, ...) imported from the unmodified files."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    from . import shims
    shims.install()

    saved = {k: v for k, v in sys.modules.items() if k == "model" or k.startswith("model.")}
    for k in saved:
        del sys.modules[k]
    # the reference's `model/` has no __init__.py (namespace package) and would lose against any
    # regular `model` package later on sys.path (this repo's mirror): hide those entries meanwhile
    saved_path = list(sys.path)
    sys.path[:] = [REFERENCE_ROOT] + [p for p in saved_path
                                      if not os.path.isdir(os.path.join(p or ".", "model")) and p != REFERENCE_ROOT]
    try:
        import importlib
        diffusion = importlib.import_module("model.diffusion")
        decoder = importlib.import_module("model.decoder")
        cmsp = importlib.import_module("model.cmsp")
        modules = importlib.import_module("model.modules")
        utils = importlib.import_module("model.utils")
        encoder = importlib.import_module("model.encoder")
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "model" or k.startswith("model.")]:
            sys.modules["_dipref_" + k] = sys.modules.pop(k)
        sys.modules.update(saved)

    ns = types.SimpleNamespace(
        diffusion=diffusion, decoder=decoder, cmsp=cmsp, modules=modules, utils=utils, encoder=encoder,
        DDIMSampler=diffusion.DDIMSampler, DDPMSampler=diffusion.DDPMSampler,
        DDPMTrainer=diffusion.DDPMTrainer, MipConditionalDecorder=decoder.MipConditionalDecorder,
        CMSP=cmsp.CMSP, root=REFERENCE_ROOT,
    )
    _cache = ns
    return ns
