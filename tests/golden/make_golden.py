"""Generate tests/golden/* from the reference source tree (run in the build
container only; /root/reference does not exist on the GPU box).

What is taken from the reference itself:
  * T_ssy_loops (code/ssy/discrete/ssy_wc_ratio.py:159-199) and T_gcy_loops
    (code/gcy/discrete/gcy_wc_ratio.py:244-302): the modules import jax and
    cannot be imported, so the single FunctionDef is extracted with ``ast`` and
    executed with numpy only.  They are fed the factor arrays of the oracle's
    discretiser (quantecon is not installable) and their outputs stored.
  * SSY()/GCY() defaults, .params and SSY.θ from the importable model files.
  * The recorded Newton trace in code/ssy/discrete/sandpit.ipynb.

usage:  python tests/golden/make_golden.py
"""
import ast
import importlib.util
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference/code"
sys.path.insert(0, ROOT)
import oracle as O  # noqa: E402


def extract(path, name):
    src = open(path, encoding="utf-8").read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"np": np}
            exec(compile(mod, path, "exec"), ns)
            return ns[name]
    raise KeyError(name)


def load_module(path, name):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    ref_T_ssy_loops = extract(f"{REF}/ssy/discrete/ssy_wc_ratio.py", "T_ssy_loops")
    ref_T_gcy_loops = extract(f"{REF}/gcy/discrete/gcy_wc_ratio.py", "T_gcy_loops")
    rng = np.random.default_rng(1233)

    ssy = O.SSY()
    for tag, shapes in (("ssy_2345", (2, 3, 4, 5)), ("ssy_4765", (4, 7, 6, 5))):
        arrays = O.discretize_ssy(ssy, shapes)
        w = np.exp(rng.standard_normal(shapes))
        w800 = np.full(shapes, 800.0)
        out = {f"arr{i}": a for i, a in enumerate(arrays)}
        out.update(shapes=np.array(shapes), params=np.array(ssy.params), w=w,
                   Tw_ref=ref_T_ssy_loops(w, shapes, ssy.params, arrays),
                   Tw800_ref=ref_T_ssy_loops(w800, shapes, ssy.params, arrays))
        np.savez_compressed(os.path.join(HERE, f"{tag}.npz"), **out)
        print(tag, "done")

    gcy = O.GCY()
    for tag, shapes in (("gcy_232323", (2, 3, 2, 3, 2, 3)),
                        ("gcy_234567", (2, 3, 4, 5, 6, 7))):
        arrays = O.discretize_gcy(gcy, shapes)
        w = np.exp(rng.standard_normal(shapes))
        out = {f"arr{i}": a for i, a in enumerate(arrays)}
        out.update(shapes=np.array(shapes), params=np.array(gcy.params), w=w,
                   Tw_ref=ref_T_gcy_loops(w, shapes, gcy.params, arrays))
        if tag == "gcy_232323":
            w800 = np.full(shapes, 800.0)
            out["Tw800_ref"] = ref_T_gcy_loops(w800, shapes, gcy.params, arrays)
        np.savez_compressed(os.path.join(HERE, f"{tag}.npz"), **out)
        print(tag, "done")

    ssy_m = load_module(f"{REF}/ssy/ssy_model.py", "ref_ssy_model")
    gcy_m = load_module(f"{REF}/gcy/gcy_model.py", "ref_gcy_model")
    rs, rg = ssy_m.SSY(), gcy_m.GCY()
    rs2 = ssy_m.SSY(γ=10.0, ψ=1.5, β=0.998)
    nb = json.load(open(f"{REF}/ssy/discrete/sandpit.ipynb"))
    trace = []
    for cell in nb["cells"]:
        for o in cell.get("outputs", []):
            for line in o.get("text", []):
                if line.startswith("iter = "):
                    trace.append(float(line.split("error = ")[1]))
    json.dump({"ssy_params": [float(x) for x in rs.params],
               "ssy_theta": float(rs.θ),
               "ssy_alt_kwargs": {"γ": 10.0, "ψ": 1.5, "β": 0.998},
               "ssy_alt_params": [float(x) for x in rs2.params],
               "ssy_alt_theta": float(rs2.θ),
               "gcy_params": [float(x) for x in rg.params],
               "sandpit_shapes": [10, 10, 10, 10],
               "sandpit_newton_errors": trace},
              open(os.path.join(HERE, "reference_facts.json"), "w"), indent=1)
    print("facts done", trace)

    # log-linear closed form (ssy_model.py:86-156, gcy_model.py:80-159), evaluated by the reference itself
    rng = np.random.default_rng(99)
    ll = {}
    for tag, mod, cls, kw, dim in (("ssy", ssy_m, ssy_m.SSY, {}, 4), ("ssy_alt", ssy_m, ssy_m.SSY, {"γ": 10.0, "ψ": 1.5, "β": 0.998}, 4),
                                   ("gcy", gcy_m, gcy_m.GCY, {}, 6), ("gcy_alt", gcy_m, gcy_m.GCY, {"γ": 9.0, "ψ": 1.8, "β": 0.998}, 6)):
        model = cls(**kw)
        f = mod.wc_loglinear_factory(model)
        pts = (rng.standard_normal((12, dim)) * ([0.01, 0.3, 0.3, 0.002] if dim == 4 else [0.01, 0.3, 0.3, 0.3, 0.002, 0.002])).tolist()
        pts.append([0.0] * dim)
        ll[tag] = {"kwargs": kw, "points": pts, "values": [float(f(tuple(x))) for x in pts]}
    json.dump(ll, open(os.path.join(HERE, "loglinear.json"), "w"), indent=1)
    print("loglinear done", ll["ssy"]["values"][-1], ll["gcy"]["values"][-1])


if __name__ == "__main__":
    main()
