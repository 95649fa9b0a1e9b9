"""Generate tests/golden/lin_interp.npz by EXECUTING the reference's interpolation utilities
(/root/reference/code/utils.py:6-23: vals_to_coords, lin_interp) -- run in the build container only.

utils.py imports jax; its two functions are pure array expressions around
``jax.scipy.ndimage.map_coordinates(fun_vals, coords, order=1, mode='nearest')``.  The FunctionDefs are
extracted with ``ast`` and executed with ``jnp`` -> numpy and ``jax.scipy.ndimage`` -> scipy.ndimage
(whose map_coordinates is the function jax's is specified against), ``jax.jit`` -> identity.

usage:  python tests/golden/make_golden_interp.py
"""
import ast
import os
import types

import numpy as np
import scipy.ndimage

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/code/utils.py"


def load_reference_utils():
    tree = ast.parse(open(REF, encoding="utf-8").read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("vals_to_coords", "lin_interp")]
    assert len(body) == 2
    jax = types.SimpleNamespace(jit=lambda f: f, scipy=types.SimpleNamespace(ndimage=scipy.ndimage))
    ns = {"jax": jax, "jnp": np}
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns["vals_to_coords"], ns["lin_interp"]


def main():
    vals_to_coords, lin_interp = load_reference_utils()
    rng = np.random.default_rng(20261018)
    out = {}
    for tag, sizes in (("d4", (4, 5, 6, 7)), ("d6", (3, 4, 3, 2, 5, 3))):
        grids = [np.linspace(-0.5 - 0.1 * i, 0.7 + 0.2 * i, n) for i, n in enumerate(sizes)]
        vals = 300.0 + 200.0 * rng.random(sizes)
        M = 400
        # inside the grid, outside it (nearest-edge extension), and exactly on nodes
        x = np.stack([rng.uniform(g[0] - 0.4 * (g[-1] - g[0]), g[-1] + 0.4 * (g[-1] - g[0]), M) for g in grids])
        for j in range(20):
            x[:, j] = [g[rng.integers(0, len(g))] for g in grids]
        y = lin_interp(x, vals, grids)
        c = vals_to_coords(grids, x)
        for i, g in enumerate(grids):
            out[f"{tag}_grid{i}"] = g
        out[f"{tag}_sizes"] = np.array(sizes)
        out[f"{tag}_vals"] = vals
        out[f"{tag}_x"] = x
        out[f"{tag}_coords_ref"] = np.asarray(c)
        out[f"{tag}_y_ref"] = np.asarray(y)
    np.savez_compressed(os.path.join(HERE, "lin_interp.npz"), **out)
    print("wrote lin_interp.npz", {k: v.shape for k, v in out.items() if k.endswith("_ref")})


if __name__ == "__main__":
    main()
