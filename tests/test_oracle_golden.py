"""Pin the numpy oracle against the reference: known-answer tests + golden vectors (CPU)."""
import numpy as np
import pytest

from oracle import jet_oracle as jo
from tests.helpers import RB2_CASES, custom_equations, load_case, rel_linf

ALL_CASES = list(RB2_CASES) + ["diffusion_leakyrelu", "ns3d_swish", "generic_d1_softplus",
                               "generic_d2_softplus", "generic_d4_softplus"]


def _equations(name, case):
    if name in RB2_CASES:
        return jo.rb2_equations(**RB2_CASES[name])
    return custom_equations(name, int(case["dim"]), case["y_f64"].shape[-1])


@pytest.mark.parametrize("d", [1, 2, 3])
def test_identity_grid_known_answer(d):
    """reference regular_nd_grid_interpolation_test.py:12-40: grid value == coordinate index."""
    axes = np.meshgrid(*([np.arange(11)] * d), indexing="ij")
    grid = np.stack(axes, axis=-1)[None].astype(np.float32)
    pts = np.random.default_rng(d).random((1, 100, d)).astype(np.float32)
    out = jo.interp(grid, pts, 0., 1., dtype=np.float32)
    np.testing.assert_allclose(out, pts * 10., atol=1e-4)


@pytest.mark.parametrize("d", [1, 2, 3])
def test_interp_coefficients_golden_bitwise(d):
    """fp32 restatement of rgi.py:14-78 reproduces the reference's tensors exactly."""
    z = np.load("tests/golden/interp_identity.npz") if False else None
    import os
    from tests.helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "interp_identity.npz"))
    cv, w, xr = jo.interp_coefficients(z[f"grid{d}"], z[f"pts{d}"], 0., 1., dtype=np.float32)
    np.testing.assert_array_equal(cv, z[f"cv{d}"])
    np.testing.assert_allclose(w, z[f"w{d}"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(xr, z[f"xr{d}"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(jo.interp(z[f"grid{d}"], z[f"pts{d}"], 0., 1., np.float32), z[f"out{d}"], atol=1e-5)


def test_heat_equation_known_answer():
    """reference pde_test.py:12-53: u=v=x^2+3y^2t+xt at (1,2,3) -> residual -7."""
    q = np.array([[[1., 2., 3.]]])
    d = 3
    env = {}
    for k, n in enumerate(("x", "y", "t")):
        g = [np.zeros((1, 1, 1)) for _ in range(d)]
        g[k] = np.ones((1, 1, 1))
        env[n] = jo.Jet(q[..., k:k + 1], g)
    x, y, t = env["x"], env["y"], env["t"]
    u = x * x + (y * y) * t * 3.0 + x * t
    env["u"] = u
    env["v"] = u
    expr = jo.parse_equation("dif(u, t) - (dif(dif(u, x), x) + dif(dif(u, y), y))", ("x", "y", "t"), ("u", "v"))
    res = jo.eval_expr_jet(expr, env, ("x", "y", "t"), d)
    np.testing.assert_allclose(u.v, 1 + 36 + 3)
    np.testing.assert_allclose(res.v, -7.0)


@pytest.mark.parametrize("name", ALL_CASES)
def test_oracle_matches_reference_fp64(name):
    c = load_case(name)
    yj = jo.query_jet(c["grid"], c["q"], c["xmin_arg"], c["xmax_arg"], c["Ws"], c["bs"], c["act"], c["act_param"])
    assert rel_linf(yj.v, c["y_f64"]) < 1e-12
    np.testing.assert_allclose(jo.query(c["grid"], c["q"], c["xmin_arg"], c["xmax_arg"], c["Ws"], c["bs"], c["act"],
                                        c["act_param"]), c["y_f64"], rtol=1e-10, atol=1e-12)
    if "g1_f64" in c:
        d = c["q"].shape[-1]
        for a in range(d):
            assert rel_linf(yj.g[a], c["g1_f64"][..., a]) < 1e-10, f"d/dq{a}"
            for b in range(d):
                assert rel_linf(yj.h[a][b], c["g2_f64"][..., a, b]) < 1e-7 or np.max(np.abs(c["g2_f64"][..., a, b])) == 0, f"d2/dq{a}dq{b}"
    in_vars, out_vars, eqs = _equations(name, c)
    res = jo.pde_residuals(yj, c["q"], in_vars, out_vars, eqs)
    for k, v in res.items():
        assert rel_linf(v, c[f"res_{k}_f64"]) < 1e-7, k


@pytest.mark.parametrize("name", ["rb2_softplus", "rb2_tanh", "rb2_leakyrelu"])
def test_reference_fp32_noise_floor_is_recorded(name):
    """The fp32 reference itself deviates from its fp64 run; the GPU gate is max(1e-5, this)."""
    c = load_case(name)
    assert rel_linf(c["y_f32"], c["y_f64"]) < 1e-5
