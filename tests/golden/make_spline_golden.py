"""Writes tests/golden/spline_basis.npz: outputs of the REFERENCE's own B-spline recursion
(`_spline_basis_vector`, /root/reference/src/lightkurve/correctors/designmatrix.py:853-893), assembled
exactly as `create_sparse_spline_matrix` does (:923-949).

`import lightkurve` fails in the build container (astropy is absent), so this script compiles that ONE
function out of the reference source file with `ast` and calls it; nothing of the reference is copied
into the repo - only its numerical outputs are frozen here.  Run from the repo root IN THE BUILD
CONTAINER (needs /root/reference): `python tests/golden/make_spline_golden.py`."""
import ast
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/src/lightkurve/correctors/designmatrix.py"


def reference_basis_function():
    mod = ast.parse(open(REF).read())
    fn = [n for n in mod.body if isinstance(n, ast.FunctionDef) and n.name == "_spline_basis_vector"][0]
    ns = {"np": np}
    exec(compile(ast.Module([fn], []), REF, "exec"), ns)
    return ns["_spline_basis_vector"]


def reference_matrix(basis, x, n_knots, degree):
    knots = np.asarray([s[-1] for s in np.array_split(np.argsort(x), n_knots - degree)[:-1]])
    knots = [np.mean([x[k], x[k + 1]]) for k in knots]
    knots = np.unique(np.append(np.append(x.min(), knots), x.max()))
    kw = np.append(np.append([x.min()] * (degree - 1), knots), [x.max()] * degree)
    cols = [basis(x, degree, idx, kw) for idx in np.arange(-1, len(kw) - degree - 1)]
    return np.asarray([c for c in cols if c.sum() != 0]).T


def main():
    basis = reference_basis_function()
    rng = np.random.default_rng(1004)
    out = {}
    cases = [(300, 20, 3, "sorted"), (257, 9, 2, "shuffled"), (120, 6, 1, "rounded"), (500, 14, 4, "sorted")]
    for i, (n, n_knots, degree, kind) in enumerate(cases):
        x = np.sort(rng.uniform(130.0, 160.0, n))
        if kind == "shuffled":
            x = rng.permutation(x)
        if kind == "rounded":                       # ties, and samples that coincide with knots
            x = np.round(x, 0)
        out["x%d" % i] = x
        out["cfg%d" % i] = np.asarray([n_knots, degree])
        out["m%d" % i] = reference_matrix(basis, x, n_knots, degree)
    np.savez_compressed(os.path.join(HERE, "spline_basis.npz"), ncases=len(cases),
                        versions="numpy %s" % np.__version__, **out)
    print("wrote spline_basis.npz")


if __name__ == "__main__":
    main()
