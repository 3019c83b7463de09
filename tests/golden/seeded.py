"""Inputs of the SEEDED golden cases: shapes too large to commit as fixtures (latent 32^3 x 128, ImNet nf=256) are
rebuilt from a seed by the generator script (tests/golden/make_golden_seeded.py, which runs the real reference on them)
and by the tests; the fixture stores the reference's outputs plus a checksum of the inputs."""
import numpy as np
import torch

CASES = {
    # name: dim, grid shape, channels, outputs, nf, activation, points, seed
    "seeded_d4_elu_nf32": dict(dim=4, gshape=(3, 3, 4, 3), c=8, o=4, nf=32, act="elu", p=64, seed=401),
    "seeded_cfg4_ns4d_nf256": dict(dim=4, gshape=(8, 8, 8, 8), c=32, o=4, nf=256, act="softplus", p=24, seed=402),
    "seeded_cfg5_rb2_nf32_c128": dict(dim=3, gshape=(32, 32, 32), c=128, o=4, nf=32, act="softplus", p=128, seed=403),
    "seeded_paper_relu": dict(dim=3, gshape=(4, 16, 16), c=32, o=4, nf=16, act="relu", p=512, seed=404),
    "seeded_paper_leakyrelu": dict(dim=3, gshape=(4, 16, 16), c=32, o=4, nf=16, act="leakyrelu", p=512, seed=405),
    "seeded_paper_elu": dict(dim=3, gshape=(4, 16, 16), c=32, o=4, nf=16, act="elu", p=512, seed=406),
}


def build(name):
    """(Ws, bs, grid [1, *gshape, c], q [1, p, dim]) as float32 torch tensors, from the case's seed only."""
    k = CASES[name]
    gen = torch.Generator().manual_seed(k["seed"])
    dimz = k["dim"] + k["c"]
    widths = [16 * k["nf"], 8 * k["nf"], 4 * k["nf"], 2 * k["nf"], k["nf"], k["o"]]
    fan_in = [dimz] + [w + dimz for w in widths[:4]] + [widths[4]]
    Ws, bs = [], []
    for n_out, n_in in zip(widths, fan_in):           # nn.Linear-like scale (uniform +- 1/sqrt(fan_in))
        bound = 1.0 / np.sqrt(n_in)
        Ws.append((torch.rand(n_out, n_in, generator=gen) * 2 - 1) * bound)
        bs.append((torch.rand(n_out, generator=gen) * 2 - 1) * bound)
    grid = torch.randn(1, *k["gshape"], k["c"], generator=gen) * 0.5
    q = torch.rand(1, k["p"], k["dim"], generator=gen) * (1 - 2e-6) + 1e-6
    return Ws, bs, grid, q


def checksum(Ws, bs, grid, q):
    return float(sum(t.double().abs().sum() for t in list(Ws) + list(bs) + [grid, q]))


def equations(name):
    """(in_vars, out_vars, [(equation name, string)]) or ("rb2", kwargs)."""
    if name.startswith("seeded_d4") or name.startswith("seeded_cfg4"):
        lap = lambda f: f"(dif(dif({f},x),x)+dif(dif({f},y),y)+dif(dif({f},z),z))"
        adv = lambda f: f"(u*dif({f},x)+v*dif({f},y)+w*dif({f},z))"
        eqs = [("mom_" + f, f"dif({f},t)+{adv(f)}+dif(p,{'xyz'['uvw'.index(f)]})-0.01*{lap(f)}") for f in "uvw"]
        eqs.append(("continuity", "dif(u,x)+dif(v,y)+dif(w,z)"))
        return "x, y, z, t", "u, v, w, p", eqs
    return "rb2", dict(t_crop=2., z_crop=1., x_crop=1., prandtl=1., rayleigh=1e6, use_continuity=True)
