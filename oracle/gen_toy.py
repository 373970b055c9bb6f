"""TEST INFRASTRUCTURE ONLY.  Turns the reference's toy instance (/root/reference/toy_example.lp: IndependentSet-15,
15 binary variables, 25 "<=" rows, 46 non-zeros; used by toy_example.py:27-60) into tests/golden/toy_is15.npz and
brute-forces its known answers (all 2^15 assignments) with plain numpy so that the feasibility / objective kernels
have a reference-owned known-answer case.  Run in the build container only (the GPU box has no /root/reference)."""
import itertools
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def parse_lp(path):
    """Minimal CPLEX-LP reader for this fixture: 'Maximize/Minimize', 'Subject to' rows 'name: +a x ... <= rhs', 'Binaries'."""
    text = open(path).read()
    text = "\n".join(l for l in text.splitlines() if not l.lstrip().startswith("\\"))
    sense = 1 if re.search(r"^\s*Maximize", text, re.M) else -1
    obj_txt = re.search(r"(?:Maximize|Minimize)\s+(.*?)Subject to", text, re.S).group(1)
    cons_txt = re.search(r"Subject to\s+(.*?)Bounds", text, re.S).group(1)
    names = re.search(r"Binaries\s+(.*?)End", text, re.S).group(1).split()
    idx = {v: i for i, v in enumerate(names)}
    term = re.compile(r"([+-]?\s*\d+(?:\.\d+)?)\s+([A-Za-z_]\w*)")
    c = np.zeros(len(names), np.int32)
    for coef, v in term.findall(obj_txt.split(":", 1)[1]):
        c[idx[v]] = int(float(coef.replace(" ", "")))
    rows, cols, vals, b = [], [], [], []
    for m in re.finditer(r"(\w+):\s*(.*?)(<=|>=|=)\s*([+-]?\d+(?:\.\d+)?)", cons_txt, re.S):
        lhs, op, rhs = m.group(2), m.group(3), float(m.group(4))
        sgn = 1 if op == "<=" else -1
        assert op != "=", "fixture has no equality rows"
        for coef, v in term.findall(lhs):
            rows.append(len(b)); cols.append(idx[v]); vals.append(sgn * int(float(coef.replace(" ", ""))))
        b.append(int(sgn * rhs))
    return dict(rows=np.array(rows, np.int64), cols=np.array(cols, np.int64), vals=np.array(vals, np.int32), b=np.array(b, np.int32),
                c=c, sense=np.int32(sense), names=np.array(names))


def main():
    d = parse_lp("/root/reference/toy_example.lp")
    n, M = len(d["c"]), len(d["b"])
    A = np.zeros((M, n), np.int64)
    np.add.at(A, (d["rows"], d["cols"]), d["vals"])
    X = np.array(list(itertools.product((0, 1), repeat=n)), np.int64)          # all 32768 assignments, row s = binary digits of s
    viol = np.maximum(X @ A.T - d["b"][None, :], 0).sum(1)
    obj = X @ d["c"].astype(np.int64)
    feas = viol == 0
    best = obj[feas].max()
    d.update(n_feasible=np.int64(feas.sum()), best_obj=np.int64(best), n_optimal=np.int64((feas & (obj == best)).sum()),
             viol_checksum=np.int64(viol.sum()), obj_checksum=np.int64((obj * feas).sum()))
    out = os.path.join(ROOT, "tests", "golden", "toy_is15.npz")
    np.savez_compressed(out, **d)
    print(out, "n", n, "M", M, "nnz", len(d["vals"]), "feasible", int(d["n_feasible"]), "best", int(best), "optimal", int(d["n_optimal"]))


if __name__ == "__main__":
    main()
