"""Instance ingestion without SCIP/ecole: a small reader for CPLEX-LP and free-format MPS files of pure binary
programs and the tensors the sampler needs from them.  Replaces, for instances that need no presolve, what the reference gets from
`extract_bigraph_from_mps` (utils.py:19-73: ecole `NodeBipartite` on the SCIP-presolved model).

What it produces (`Instance`): every row in "<=" form with integer coefficients (`>=` rows are negated, `=` rows become two rows),
`c` for MINIMISATION (a Maximize objective is negated), the float tensors in the conventions of dip_b200.synth -- `b` divided by
the row's coefficient norm, `c` by the objective norm, column 0 of the constraint / variable feature matrices holding b / c as the
sampler reads them (sample.py:146-148) -- and the raw integers for the exact feasibility check.  The remaining ecole feature
columns depend on an LP solve and on SCIP's internal state; they are filled with static stand-ins (row degree, objective cosine,
variable degree, objective rank, ...) and documented as such: a model trained on real ecole features needs the real features.
"""
import math
import re
from types import SimpleNamespace

import numpy as np

_NUM = r"(?:\d+(?:\.\d*)?|\.\d+)(?:[eE][+-]?\d+)?"          # 3, 3., 3.5, .5, 1e3
_TERM = re.compile(r"([+-]?)\s*(" + _NUM + r")?\s*\*?\s*([A-Za-z_!\"#$%&()/,;?@'`{}|~][\w!\"#$%&()/,.;?@'`{}|~\[\]]*)")
_ROW = re.compile(r"(?:([A-Za-z_][\w.\[\]]*)\s*:)?\s*([^:<>=]*?)\s*(<=|=<|>=|=>|<|>|=)\s*([+-]?\s*" + _NUM + r")")
_SECTIONS = ("maximize", "maximise", "max", "minimize", "minimise", "min", "subject to", "such that", "st", "s.t.", "bounds", "bound",
             "binaries", "binary", "bin", "generals", "general", "gen", "end")


def _terms(expr):
    out = []
    for sign, coef, name in _TERM.findall(expr):
        v = float(coef) if coef else 1.0
        out.append((name, -v if sign == "-" else v))
    return out


def parse_lp(text):
    """CPLEX-LP subset: objective, linear rows with <=, >=, =, a Bounds section (only 0/1 bounds accepted), Binaries."""
    lines = [l.split("\\")[0].rstrip() for l in text.splitlines()]
    sections, cur = {}, None
    for l in lines:
        key = l.strip().lower()
        if key in _SECTIONS:
            cur = {"maximise": "maximize", "max": "maximize", "minimise": "minimize", "min": "minimize", "such that": "subject to", "st": "subject to",
                   "s.t.": "subject to", "bound": "bounds", "binary": "binaries", "bin": "binaries", "general": "generals", "gen": "generals"}.get(key, key)
            sections.setdefault(cur, [])
        elif cur is not None and l.strip():
            sections[cur].append(l.strip())
    if "generals" in sections and sections["generals"]:
        raise ValueError("general integer variables are not supported (binary programs only)")
    sense = -1 if "maximize" in sections else 1
    obj_txt = " ".join(sections.get("maximize", sections.get("minimize", [])))
    obj_txt = obj_txt.split(":", 1)[1] if ":" in obj_txt else obj_txt
    names = " ".join(sections.get("binaries", [])).split()
    if not names:
        raise ValueError("no Binaries section: binary programs only")
    index = {v: i for i, v in enumerate(names)}
    c = np.zeros(len(names), dtype=np.int64)

    def col(nm, where):
        if nm not in index:
            raise ValueError("%s uses '%s', which is not declared in the Binaries section (binary programs only)" % (where, nm))
        return index[nm]

    for nm, v in _terms(obj_txt):
        if v != round(v):
            raise ValueError("objective coefficient %g of %s is not an integer (the exact objective check works on integers)" % (v, nm))
        c[col(nm, "the objective")] += int(round(sense * v))
    for b in sections.get("bounds", []):
        if not re.fullmatch(r"\s*(?:0\s*<=\s*)?\S+\s*<=\s*1\s*|\s*\S+\s*>=\s*0\s*|\s*0\s*<=\s*\S+\s*", b):
            raise ValueError("only 0/1 bounds are supported (got '%s')" % b)
    rows = []
    cons_txt = " ".join(sections.get("subject to", []))
    # every row is "[label:] expr op rhs": rows are delimited by their relational operator + right-hand side, so unlabeled rows
    # are read one by one like labeled ones; anything left between two rows is an error, not silently dropped
    pos = 0
    for k, m in enumerate(_ROW.finditer(cons_txt)):
        if cons_txt[pos:m.start()].strip():
            raise ValueError("cannot parse '%s' in the constraint section" % cons_txt[pos:m.start()].strip())
        label = m.group(1) or ("r%d" % k)
        op = {"<": "<=", ">": ">=", "=<": "<=", "=>": ">="}.get(m.group(3), m.group(3))
        body = m.group(2)
        terms = _terms(body)
        if not terms or _TERM.sub("", body).strip(" +-"):
            raise ValueError("row %s: cannot parse '%s'" % (label, body))
        rows.append((label, [(col(nm, "row " + label), v) for nm, v in terms], op, float(m.group(4).replace(" ", ""))))
        pos = m.end()
    if cons_txt[pos:].strip():
        raise ValueError("cannot parse '%s' at the end of the constraint section" % cons_txt[pos:].strip())
    return _assemble(names, c, rows, sense)


def parse_mps(text):
    """Free-format MPS of a binary program: ROWS (N, L, G, E), COLUMNS (MARKER lines ignored), RHS, BOUNDS (BV / UP 1 / LO 0)."""
    sec, rows_kind, order, cols, rhs, names = None, {}, [], {}, {}, []
    obj_row, sense = None, 1
    binary, marked, in_int = set(), set(), False
    for raw in text.splitlines():
        if not raw.strip() or raw.lstrip().startswith("*"):
            continue
        tok = raw.split()
        if not raw[0].isspace():
            sec = tok[0].upper()
            continue
        if sec == "OBJSENSE":
            sense = -1 if tok[0].upper().startswith("MAX") else 1
        elif sec == "ROWS":
            kind, nm = tok[0].upper(), tok[1]
            if kind == "N" and obj_row is None:
                obj_row = nm
            else:
                rows_kind[nm] = kind
                order.append(nm)
        elif sec == "COLUMNS":
            if len(tok) >= 3 and tok[1].upper() == "'MARKER'":
                in_int = "INTORG" in tok[2].upper()
                continue
            col = tok[0]
            if col not in cols:
                cols[col] = {}
                names.append(col)
            if in_int:
                marked.add(col)
            for r, v in zip(tok[1::2], tok[2::2]):
                cols[col][r] = cols[col].get(r, 0.0) + float(v)
        elif sec == "RHS":
            for r, v in zip(tok[1::2], tok[2::2]):
                rhs[r] = float(v)
        elif sec == "BOUNDS":
            kind = tok[0].upper()
            val = float(tok[3]) if len(tok) > 3 else None
            if not (kind == "BV" or (kind == "UP" and val == 1.0) or (kind == "LO" and val == 0.0)):
                raise ValueError("only 0/1 bounds are supported (got %s)" % raw.strip())
            if kind in ("BV", "UP"):
                binary.add(tok[2])
        elif sec == "RANGES":
            raise ValueError("a RANGES section is not supported (write ranged rows as two rows)")
        elif sec not in ("NAME", "OBJSENSE", "ROWS", "COLUMNS", "RHS", "BOUNDS", "ENDATA", "OBJSENSE:"):
            raise ValueError("MPS section %s is not supported" % sec)
    loose = [v for v in names if v not in binary]
    if loose:
        # a column without a BV / "UP 1" bound is continuous (or, inside integer markers, a general integer) in MPS: not a binary program
        raise ValueError("columns without a 0/1 bound (binary programs only): %s%s" % (", ".join(loose[:5]), " ..." if len(loose) > 5 else ""))
    index = {v: i for i, v in enumerate(names)}
    c = np.zeros(len(names), dtype=np.int64)
    per_row = {r: [] for r in order}
    for col, ent in cols.items():
        for r, v in ent.items():
            if r == obj_row:
                if v != round(v):
                    raise ValueError("objective coefficient %g of %s is not an integer (the exact objective check works on integers)" % (v, col))
                c[index[col]] += int(round(sense * v))
            else:
                per_row[r].append((index[col], v))
    rows = [(r, per_row[r], {"L": "<=", "G": ">=", "E": "="}[rows_kind[r]], rhs.get(r, 0.0)) for r in order]
    return _assemble(names, c, rows, sense)


def _assemble(names, c_int, rows, sense):
    ri, ci, vi, bi, labels = [], [], [], [], []
    for label, coefs, op, rhs in rows:
        variants = {"<=": [1], "=<": [1], ">=": [-1], "=>": [-1], "=": [1, -1]}[op]
        for sg in variants:
            merged = {}
            for k, v in coefs:
                merged[k] = merged.get(k, 0.0) + sg * v
            for k in sorted(merged):
                if merged[k] != 0.0:
                    if merged[k] != round(merged[k]):
                        raise ValueError("row %s: non-integer coefficient" % label)
                    ri.append(len(bi)); ci.append(k); vi.append(int(round(merged[k])))
            if sg * rhs != round(sg * rhs):
                raise ValueError("row %s: non-integer right-hand side" % label)
            bi.append(int(round(sg * rhs)))
            labels.append(label)
    return make_instance(names, np.array(ri, np.int64), np.array(ci, np.int64), np.array(vi, np.int32), np.array(bi, np.int32),
                         c_int.astype(np.int32), sense, labels)


def make_instance(names, rows, cols, a_int, b_int, c_int, sense=1, row_labels=None):
    n, M = len(names), len(b_int)
    row_norm = np.sqrt(np.bincount(rows, weights=a_int.astype(np.float64) ** 2, minlength=M))
    row_norm[row_norm == 0] = 1.0
    a_val = (a_int / row_norm[rows]).astype(np.float32)
    b = (b_int / row_norm).astype(np.float32)
    c_norm = float(np.linalg.norm(c_int)) or 1.0
    c = (c_int / c_norm).astype(np.float32)
    cons = np.zeros((M, 5), dtype=np.float32)
    var = np.zeros((n, 17), dtype=np.float32)
    cons[:, 0] = b
    var[:, 0] = c
    # static stand-ins for the LP-dependent ecole columns (see the module docstring)
    deg_r = np.bincount(rows, minlength=M).astype(np.float32)
    deg_c = np.bincount(cols, minlength=n).astype(np.float32)
    dot = np.bincount(rows, weights=a_int * c_int[cols].astype(np.float64), minlength=M)
    cons[:, 1] = (dot / (row_norm * c_norm)).astype(np.float32)           # cosine between the row and the objective
    cons[:, 2] = deg_r / max(n, 1)
    var[:, 1] = 1.0                                                        # binary type indicator
    var[:, 2] = deg_c / max(M, 1)
    var[:, 3] = (np.argsort(np.argsort(c_int)) / max(n - 1, 1)).astype(np.float32)
    return SimpleNamespace(
        kind="file", M=M, N=n, n=n, nnz=len(rows), rows=rows, cols=cols, a_val=a_val, b=b, c=c, a_int=a_int, b_int=b_int, c_int=c_int,
        constraint_features=cons, variable_features=var, edge_index=np.stack([rows, cols]).astype(np.int64), edge_attr=a_val[:, None].copy(),
        int_indices=np.arange(n, dtype=np.int64), n_int_vars=n, var_names=list(names), sense=sense, row_labels=row_labels or [])


def write_lp(inst, name="instance"):
    """The instance back as CPLEX-LP text (minimisation form, every row '<=')."""
    def expr(idx, vals):
        return " ".join("%+d %s" % (int(v), inst.var_names[k]) for k, v in zip(idx, vals)) or "0 %s" % inst.var_names[0]
    out = ["\\ %s" % name, "Minimize", " obj: " + expr(range(inst.n), inst.c_int), "Subject to"]
    order = np.argsort(inst.rows, kind="stable")
    ptr = np.searchsorted(inst.rows[order], np.arange(inst.M + 1))
    for i in range(inst.M):
        e = order[ptr[i]:ptr[i + 1]]
        out.append(" c_%d: %s <= %+d" % (i, expr(inst.cols[e], inst.a_int[e]), int(inst.b_int[i])))
    out += ["Bounds"] + [" 0 <= %s <= 1" % v for v in inst.var_names] + ["Binaries", " " + " ".join(inst.var_names), "End"]
    return "\n".join(out) + "\n"


def read(path):
    text = open(path).read()
    return parse_mps(text) if path.lower().endswith(".mps") else parse_lp(text)
