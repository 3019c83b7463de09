"""GPU parity (-m gpu) on the SEEDED golden cases (tests/golden/seeded.py): the shapes whose tolerances used to be relaxed
past 1e-5 against the oracle only - d = 4, BASELINE config 4 (ImNet nf=256, 8^4 x 32 latent), config 5 (32^3 x 128
latent, nf=32) and the kinked activations at the paper shape.  Every gate is

    max(1e-5, 2 * rel-Linf(reference float32, reference float64))        (rel-Linf = max|a-b| / max|b|, plain L-infinity)

with both reference runs stored in the fixture by tests/golden/make_golden_seeded.py (real reference, this container)."""
import os
import sys

import numpy as np
import pytest
import torch

import space_time_pde_b200 as sp
from space_time_pde_b200 import jets
from space_time_pde_b200.equations import JetSpec
from tests.helpers import GOLDEN, record, rel_err_quantile, rel_linf

sys.path.insert(0, GOLDEN)
import seeded  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need CUDA (the hot path has no CPU fallback)"
    return torch.device("cuda:0")


def make_layer(name):
    eq = seeded.equations(name)
    if eq[0] == "rb2":
        return sp.get_rb2_pde_layer(**eq[1])
    layer = sp.PDELayer(in_vars=eq[0], out_vars=eq[1])
    for eq_name, string in eq[2]:
        layer.add_equation(string, eq_name)
    return layer


def gate_of(z, key):
    ref = z[key + "_f64"]
    return max(1e-5, 2 * rel_linf(z[key + "_f32"], ref)), ref


@pytest.mark.parametrize("precision", ["fp16x3", "fp32"])
@pytest.mark.parametrize("name", list(seeded.CASES))
def test_seeded_golden(name, precision, dev):
    k = seeded.CASES[name]
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    Ws, bs, grid, q = seeded.build(name)
    assert abs(seeded.checksum(Ws, bs, grid, q) - float(z["checksum"])) < 1e-9 * float(z["checksum"]), \
        "seeded inputs differ from the ones the reference ran on (torch RNG changed?): regenerate the fixture"
    model = sp.ImNet(dim=k["dim"], in_features=k["c"], out_features=k["o"], nf=k["nf"],
                     activation=sp.NONLINEARITIES[k["act"]])
    with torch.no_grad():
        for i in range(6):
            model.fc[i].weight.copy_(Ws[i])
            model.fc[i].bias.copy_(bs[i])
    model, grid, q = model.to(dev), grid.to(dev), q.to(dev)
    layer = make_layer(name)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
    d = k["dim"]
    spec = JetSpec(tuple(range(d)), tuple((a, a) for a in range(d)))
    jets.set_default_precision(precision)
    try:
        with torch.no_grad():
            y, res = layer(q)
            y2, jt = sp.fused_query(grid, q, 0., 1., list(model.fc), k["act"], None, spec=spec)
    finally:
        jets.set_default_precision("fp16x3")
    tag = f"{name}[{precision}]"
    kinked = k["act"] in ("relu", "leakyrelu")

    def check(what, got, gate, ref):
        """L-infinity against the measured-noise gate.  relu / leakyrelu: a pre-activation within rounding distance of 0
        flips sigma' between its two values at isolated points, and WHICH points flip depends on the summation order - the
        reference's float32 run flips at its own points (2.2e-5 on the leakyrelu fixture), this implementation at others
        (measured 5.9e-5 / 7.4e-5 on the relu fixture, in the tensor-core mode AND in the plain-FFMA fp32 mode).  So for
        those two activations the noise gate applies to the 99.5 % quantile over points and the L-infinity, recorded
        beside it, is bounded by 1e-3 (1e-2 for second derivatives: measured 1.8e-3 on d1d1 of the relu fixture)."""
        linf = record(tag, what, rel_linf(got, ref), gate)
        if kinked:
            assert rel_err_quantile(got, ref) < gate, what
            assert linf < (1e-3 if len(what) < 4 else 1e-2), what   # second derivatives (dXdX) of a flipped unit: 1e-2
        else:
            assert linf < gate, what

    gate, ref = gate_of(z, "y")
    check("y", y.cpu().numpy(), gate, ref)
    for key, v in res.items():
        gate, ref = gate_of(z, "res_" + key)
        check("res_" + key, v.cpu().numpy(), gate, ref)
    jt = jt.cpu().numpy()
    for a in range(d):
        ref1 = z["g1_f64"][..., a]
        g1 = max(1e-5, 2 * rel_linf(z["g1_f32"][..., a], ref1))
        check(f"d{a}", jt[spec.plane((a,))], g1, ref1)
        ref2 = z["g2diag_f64"][..., a]
        if np.max(np.abs(ref2)) == 0:                      # relu family: sigma'' = 0 everywhere
            assert np.max(np.abs(jt[spec.plane((a, a))])) == 0
            continue
        g2 = max(1e-5, 2 * rel_linf(z["g2diag_f32"][..., a], ref2))
        check(f"d{a}d{a}", jt[spec.plane((a, a))], g2, ref2)


def test_reference_noise_table():
    """The measured reference-float32 noise that justifies every gate above 1e-5 (printed with -s, recorded in the report)."""
    for name in seeded.CASES:
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        for key in z.files:
            if key.endswith("_f32"):
                n = rel_linf(z[key], z[key[:-4] + "_f64"])
                record("reference_fp32_noise", f"{name}:{key[:-4]}", n, 0.0)
                print(f"{name:30s} {key[:-4]:24s} reference fp32 vs fp64 rel-Linf {n:.1e}")
