"""Shared helpers for the parity tests (fixtures -> oracle / product inputs)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

RB2_CASES = {
    "rb2_tanh": dict(t_crop=2., z_crop=1., x_crop=1., use_continuity=True),
    "rb2_relu": dict(t_crop=2., z_crop=1., x_crop=1., use_continuity=True),
    "rb2_softplus": dict(t_crop=2., z_crop=1., x_crop=1., use_continuity=True),
    "rb2_elu": dict(t_crop=2., z_crop=1., x_crop=1., use_continuity=True),
    "rb2_swish": dict(t_crop=2., z_crop=1., x_crop=1., use_continuity=True),
    "rb2_leakyrelu": dict(t_crop=2., z_crop=1., x_crop=1., use_continuity=True),
    "rb2_paper_softplus": dict(mean=[0.1, -0.2, 0.05, 0.3], std=[1.1, 0.9, 1.3, 0.7], t_crop=2., z_crop=1.,
                               x_crop=2., use_continuity=True),
    "rb2_nonunit_tanh": dict(use_continuity=False),
    "rb2_ties_softplus": dict(use_continuity=True),
}


def load_case(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    case = {k: z[k] for k in z.files}
    case["act"] = str(case["act"])
    case["Ws"] = [case[f"W{i}"] for i in range(6)]
    case["bs"] = [case[f"b{i}"] for i in range(6)]
    case["act_param"] = float(case.get("act_param", 1.0))
    xmax = case["xmax"]
    case["xmax_arg"] = float(xmax) if xmax.ndim == 0 else xmax.astype(np.float32)
    case["xmin_arg"] = 0.0 if xmax.ndim == 0 else np.zeros_like(case["xmax_arg"])
    return case


def custom_equations(name, dim=None, o=None):
    """(in_vars, out_vars, {eq_name: (string, subs)}) for the non-RB2 golden cases."""
    if name.startswith("diffusion"):
        return ("x", "y", "t"), ("u", "v"), {
            "diffusion_u": ("dif(u, t) - (dif(dif(u, x), x) + dif(dif(u, y), y))", None),
            "diffusion_v": ("dif(v, t) - (dif(dif(v, x), x) + dif(dif(v, y), y))", None)}
    if name.startswith("ns3d"):
        lap = lambda f: f"(dif(dif({f},x),x)+dif(dif({f},y),y)+dif(dif({f},z),z))"
        adv = lambda f: f"(u*dif({f},x)+v*dif({f},y)+w*dif({f},z))"
        return ("x", "y", "z"), ("u", "v", "w", "p"), {
            "mom_u": (f"{adv('u')}+dif(p,x)-0.01*{lap('u')}", None),
            "mom_v": (f"{adv('v')}+dif(p,y)-0.01*{lap('v')}", None),
            "mom_w": (f"{adv('w')}+dif(p,z)-0.01*{lap('w')}", None),
            "continuity": ("dif(u,x)+dif(v,y)+dif(w,z)", None)}
    if name.startswith("generic"):
        names_in = ["x", "y", "z", "s"][:dim]
        names_out = ["u", "v", "w", "r"][:o]
        a, b = names_in[0], names_in[-1]
        return tuple(names_in), tuple(names_out), {
            "mixed": (f"dif(dif({names_out[0]},{a}),{b}) + {a}*dif({names_out[-1]}*{names_out[0]},{b})", None)}
    raise KeyError(name)


def rel_linf(a, b):
    """max|a-b| / max|b| (the north star's rel-L-infinity)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


def rel_err_quantile(a, b, q=0.995):
    """q-quantile over points of |a-b| / max|b| - for activations with kinks (relu family) a handful of
    points whose pre-activation sits within rounding distance of 0 flip sigma' between 0 and 1; the
    bulk of the points must still agree to tolerance."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    e = np.abs(a - b) / (den if den > 0 else 1.0)
    qv = float(np.quantile(e, q))
    # the plain L-infinity is reported beside every quantile-gated comparison (parity report, see record())
    record("linf_beside_quantile", os.environ.get("PYTEST_CURRENT_TEST", "?").split(" ")[0], float(np.max(e)), qv)
    return qv


GRAD_CASES = ["rb2_tanh", "rb2_softplus", "rb2_elu", "rb2_swish", "rb2_paper_softplus", "rb2_nonunit_tanh",
              "generic_d1_softplus", "generic_d2_softplus", "generic_d4_softplus"]


def load_grads(name):
    """Gradients of the training-style loss from the REAL reference (tests/golden/make_golden_grads.py, float64)."""
    z = np.load(os.path.join(GOLDEN, "grads_" + name + ".npz"))
    return {k: z[k] for k in z.files}


def pde_layer_for(sp, name, case):
    """The PDELayer of a golden case built on the product's API (RB2 or the custom strings)."""
    if name in RB2_CASES:
        return sp.get_rb2_pde_layer(**RB2_CASES[name])
    dim, o = int(case["dim"]), case["Ws"][5].shape[0]
    in_vars, out_vars, eqs = custom_equations(name, dim, o)
    layer = sp.PDELayer(in_vars=", ".join(in_vars), out_vars=", ".join(out_vars))
    for eq_name, (string, subs) in eqs.items():
        layer.add_equation(string, eq_name, subs_dict=subs)
    return layer


def record(test, what, err, gate):
    """Append one measured parity figure to the JSON-lines file named by STPDE_PARITY_REPORT (the GPU sessions set it;
    the committed copy is profiles/r02_parity_report.jsonl).  Returns err so that it can sit inside an assert."""
    path = os.environ.get("STPDE_PARITY_REPORT")
    if path:
        import json
        with open(path, "a") as f:
            f.write(json.dumps({"test": test, "what": what, "err": float(err), "gate": float(gate)}) + "\n")
    return err
