"""Partial-solution hand-off to the host solver (reference: sample_partialsol.py:164-199).

After sampling, the reference turns every decoded 0/1 assignment into a SCIP *partial solution*: `int(n * coverage)` variables are
drawn with Python's `random.sample(range(n), ...)` and only those get a value (`scip_model.setSolVal`), the rest is left to the
`completesol` heuristic.  SCIP itself stays host-side glue outside this package; what this module produces is exactly what that
glue consumes -- per sample the {variable name: 0/1} map -- plus a compact bit-packed form (selection mask + values) for moving
10 000-instance batches between ranks or to solver workers.

Reference quirk, kept by default: the value written for sample j is `pred_x[0, k]` (sample_partialsol.py:179 indexes sample 0 for
every j), so all partial solutions of an instance carry the FIRST sample's values and differ only in which variables are fixed.
`reference_quirk=False` uses sample j's own assignment.
"""
import random as _random

import numpy as np


def select_variables(n, coverage, rng=None):
    """`random.sample(range(n), int(n * coverage))` with the module-level RNG of the reference (or an explicit random.Random)."""
    rng = _random if rng is None else rng
    return rng.sample(range(n), int(n * coverage))


def partial_solutions(xbits, var_names, coverage, rng=None, reference_quirk=True):
    """xbits: (B, n) array of 0/1 (device tensors are copied to the host); var_names: n names in the column order of A (the
    reference strips SCIP's 't_' prefix, sample_partialsol.py:179).  Returns a list of B dicts {name: int} in sample order; the
    RNG is advanced once per sample exactly like the reference's loop."""
    x = np.asarray(xbits.cpu() if hasattr(xbits, "cpu") else xbits)
    if x.ndim != 2 or x.shape[1] != len(var_names):
        raise ValueError("xbits must be (B, n) with n == len(var_names)")
    B, n = x.shape
    out = []
    for j in range(B):
        src = x[0] if reference_quirk else x[j]
        chosen = select_variables(n, coverage, rng)
        out.append({var_names[k]: int(src[k]) for k in sorted(chosen)})
    return out


def pack(xbits, selections):
    """Compact form: (B, ceil(n/8)) uint8 value bits and selection-mask bits (numpy.packbits, big-endian within a byte)."""
    x = np.asarray(xbits.cpu() if hasattr(xbits, "cpu") else xbits).astype(np.uint8)
    B, n = x.shape
    mask = np.zeros((B, n), dtype=np.uint8)
    for j, sel in enumerate(selections):
        mask[j, np.asarray(sel, dtype=np.int64)] = 1
    return np.packbits(x & mask, axis=1), np.packbits(mask, axis=1)


def unpack(values, mask, n):
    """Inverse of pack: list of {column index: value} per sample."""
    v = np.unpackbits(values, axis=1)[:, :n]
    m = np.unpackbits(mask, axis=1)[:, :n]
    return [{int(k): int(v[j, k]) for k in np.flatnonzero(m[j])} for j in range(v.shape[0])]


# ---- device-side path ---------------------------------------------------------------------------
def select_gpu(xbits, coverage, seed=0, instance_id=0, sample_base=0, reference_quirk=True):
    """Selection + bit-packing on the GPU (dip_partial_select): xbits (B, n) uint8 CUDA tensor as returned by Engine.final_decode /
    sample_guided_ddim.  Returns (values, mask) device tensors in the format of `pack`.  The subset of sample j is a pure function of
    (seed, instance_id, sample_base + j) -- any sharding of samples over ranks selects the same variables -- and has exactly
    int(n * coverage) members like the reference's `random.sample(range(n), num_vars)` (sample_partialsol.py:171,190)."""
    from . import ops
    n = xbits.shape[1]
    return ops.partial_select(xbits, int(n * coverage), seed, instance_id, sample_base, first_sample_values=reference_quirk)


def to_sol_text(var_names, values_row, mask_row, objective=None):
    """One sample's partial solution in SCIP's .sol text format (name value per line, only the selected variables, in the column
    order of A), ready for `readSolFile` / `scip -c "read x.sol"`; values_row / mask_row: one row of the packed form."""
    n = len(var_names)
    v = np.unpackbits(np.asarray(values_row, dtype=np.uint8))[:n]
    m = np.unpackbits(np.asarray(mask_row, dtype=np.uint8))[:n]
    lines = ["solution status: unknown"]
    if objective is not None:
        lines.append("objective value: %s" % objective)
    lines += ["%s %d" % (var_names[k], int(v[k])) for k in np.flatnonzero(m)]
    return "\n".join(lines) + "\n"
