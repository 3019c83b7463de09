"""CPU tests (-m "not gpu"): library exports, descriptor layout, equation compiler, reference API
behaviour.  The numeric host-logic tests install the torch-op jet evaluator as a stand-in for the
CUDA kernel (space_time_pde_b200.jets.set_test_backend) - the product never does that."""
import ctypes
import os
import re

import numpy as np
import pytest
import sympy
import torch

import space_time_pde_b200 as sp
from space_time_pde_b200 import _lib, _torch_jets, equations, jets
from tests.helpers import GRAD_CASES, RB2_CASES, custom_equations, load_case, load_grads, pde_layer_for, rel_linf

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture()
def cpu_backend():
    def backend(grid, q, lo, hi, Ws, bs, act, beta, spec):
        return _torch_jets.query_jets(grid, q, lo, hi, list(Ws), list(bs), act, torch.tensor(beta), spec)
    jets.set_test_backend(backend)
    yield
    jets.set_test_backend(None)


def build_model(c, out_features):
    act = c["act"]
    model = sp.ImNet(dim=int(c["dim"]), in_features=c["grid"].shape[-1], out_features=out_features, nf=int(c["nf"]),
                     activation=sp.NONLINEARITIES[act])
    with torch.no_grad():
        for i in range(6):
            model.fc[i].weight.copy_(torch.tensor(c["Ws"][i]))
            model.fc[i].bias.copy_(torch.tensor(c["bs"][i]))
        if act == "swish":
            model.activ.beta.fill_(c["act_param"])
    return model


def bounds(c):
    if np.isscalar(c["xmax_arg"]):
        return 0., c["xmax_arg"]
    return torch.zeros(len(c["xmax_arg"])), torch.tensor(c["xmax_arg"])


# ---------------------------------------------------------------------------------------------
# C ABI: the library loads and exports every symbol include/stpde.h declares
# ---------------------------------------------------------------------------------------------
def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "stpde.h")).read()
    declared = set(re.findall(r"\b(stpde_[a-z_0-9]+)\s*\(", header))
    declared -= {"stpde_desc"}
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), f"libstpde.so does not export {name}"
    assert set(_lib.EXPORTS) == declared
    # every symbol of the shared object resolves at load time (nvcc -shared links with undefined symbols allowed)
    ctypes.CDLL(_lib.LIB_PATH, mode=os.RTLD_NOW)
    assert lib.stpde_version() == 200
    assert lib.stpde_desc_size() == ctypes.sizeof(_lib.StpdeDesc)


def test_workspace_bytes_and_descriptor_validation():
    lib = _lib.load()
    d = _lib.StpdeDesc()
    assert lib.stpde_workspace_bytes(ctypes.byref(d)) == 0          # dim = 0 is invalid
    assert b"dim" in lib.stpde_last_error()
    grid = torch.zeros(1, 4, 16, 16, 32)
    q = torch.zeros(1, 4096, 3)
    lo, hi = jets.bounds_tensors(0., 1., 3, "cpu")
    spec = equations.JetSpec((0, 1, 2), ((1, 1), (2, 2)))
    d = jets.make_desc(grid, q, lo, hi, [512, 256, 128, 64, 32, 4], "softplus", 1.0, spec, "fp32")
    n = lib.stpde_workspace_bytes(ctypes.byref(d))
    assert n > 4096 * 8 * 6 * (512 + 256) * 4                         # activations of one chunk
    d.n_second = 11
    assert lib.stpde_workspace_bytes(ctypes.byref(d)) == 0


def test_hot_path_fails_loudly_without_cuda():
    model = sp.ImNet(dim=3, in_features=8, out_features=4, nf=4)
    with pytest.raises(RuntimeError, match="CUDA"):
        sp.query_local_implicit_grid(model, torch.rand(1, 3, 3, 3, 8), torch.rand(1, 5, 3), 0., 1.)
    with pytest.raises(RuntimeError, match="CUDA"):
        sp.regular_nd_grid_interpolation(torch.rand(1, 3, 3, 2), torch.rand(1, 5, 2), 0., 1.)


# ---------------------------------------------------------------------------------------------
# equation compiler
# ---------------------------------------------------------------------------------------------
def test_rb2_jet_spec_is_six_components():
    layer = sp.get_rb2_pde_layer(t_crop=2., z_crop=1., x_crop=1., use_continuity=True)
    spec = layer.jet_spec()
    assert spec.first == (0, 1, 2)
    assert spec.second == ((1, 1), (2, 2))          # x and z bind to columns 1 and 2 (quirk Q4)
    assert layer.eqn_names == ["transport_eqn_b", "transport_eqn_u", "transport_eqn_w", "continuity"]
    _, program = layer._binding()
    assert program is not None and len(program[0]) % 2 == 0


def test_normalised_equations_chain_rule():
    layer = sp.get_rb2_pde_layer(mean=[0.1, -0.2, 0.05, 0.3], std=[1.1, 0.9, 1.3, 0.7], use_continuity=True)
    ce = layer.eqns_jet["continuity"]
    # d(u*1.3+0.05)/dx = 1.3 u_x ;  0.5 * 1.3 = 0.65
    coeff = ce.expr.coeff(sympy.Symbol("u__x"))
    assert abs(float(coeff) - 0.65) < 1e-12


def test_third_derivative_falls_back_to_autograd_route():
    layer = sp.PDELayer("x, y", "u")
    layer.add_equation("dif(dif(dif(u,x),x),x)", "third")
    assert layer.eqns_jet["third"] is None and layer.jet_spec() is None


def test_jet_spec_split():
    full = equations.JetSpec((0, 1, 2, 3), tuple((a, b) for a in range(4) for b in range(a, 4)))
    parts = full.split(10)
    assert all(1 + len(p.first) + len(p.second) <= 10 for p in parts)
    assert sorted(sum((list(p.second) for p in parts), [])) == sorted(full.second)


# ---------------------------------------------------------------------------------------------
# reference API / error behaviour (src/pde.py, quirk Q5) and the reference's own tests
# ---------------------------------------------------------------------------------------------
def test_pde_layer_error_behaviour():
    layer = sp.PDELayer(in_vars="x, y", out_vars="u")
    with pytest.raises(KeyError):
        layer.add_equation("dif(u,x)")                      # reference pde.py:64
    with pytest.raises(ValueError):
        layer.add_equation("dif(u,x)+q", "bad")             # unknown symbol, pde.py:74-79
    layer.add_equation("dif(u,x)", "ok")
    with pytest.raises(RuntimeError):
        layer(torch.zeros(1, 2))                             # no forward method, pde.py:105-107
    layer.update_forward_method(lambda x: torch.cat([x, x], dim=-1))
    with pytest.raises(ValueError):
        layer(torch.zeros(1, 2))                             # output dims, pde.py:109-112
    assert layer.eqn_num == 1 and layer.n_in == 2 and layer.n_out == 1


def test_heat_equation_known_answer_generic_forward():
    """reference src/pde_test.py:12-53 (autograd route: the forward is an arbitrary torch function)."""
    def fwd_fn(inpt):
        u = inpt[..., 0:1]**2 + 3*inpt[..., 1:2]**2*inpt[..., 2:3] + inpt[..., 0:1]*inpt[..., 2:3]
        return torch.cat([u, u], axis=-1)
    layer = sp.PDELayer(in_vars='x, y, t', out_vars='u, v')
    layer.add_equation('dif(u, t) - (dif(dif(u, x), x) + dif(dif(u, y), y))', 'diffusion_u')
    layer.add_equation('dif(v, t) - (dif(dif(v, x), x) + dif(dif(v, y), y))', 'diffusion_v')
    layer.update_forward_method(fwd_fn)
    inpt = torch.tensor([[1., 2., 3.]])
    val, grads = layer(inpt)
    np.testing.assert_allclose(val.detach().numpy(), fwd_fn(inpt).numpy(), atol=1e-4)
    for name in ('diffusion_u', 'diffusion_v'):
        np.testing.assert_allclose(grads[name].detach().numpy(), [[-7.0]])


def test_imnet_shapes_and_state_dict_keys():
    """reference src/implicit_net_test.py:15-26 + checkpoint key compatibility."""
    model = sp.ImNet(dim=4, in_features=32, out_features=3, nf=16)
    out = model(torch.rand(32 * 64, 36))
    assert tuple(out.shape) == (32 * 64, 3)
    keys = set(model.state_dict().keys())
    for i in range(6):
        assert {f"fc{i}.weight", f"fc{i}.bias", f"fc.{i}.weight", f"fc.{i}.bias"} <= keys
    sw = sp.ImNet(activation=sp.NONLINEARITIES["swish"])
    assert "activ.beta" in sw.state_dict()
    assert [l.in_features for l in model.fc] == [36, 256 + 36, 128 + 36, 64 + 36, 32 + 36, 16]


def test_decoder_signature_duck_typing():
    from space_time_pde_b200.implicit_net import decoder_signature
    model = sp.ImNet(dim=3, in_features=8, out_features=4, nf=4, activation=torch.nn.Softplus)
    layers, act, param = decoder_signature(torch.nn.DataParallel(model))
    assert act == "softplus" and param is None and len(layers) == 6
    assert decoder_signature(torch.nn.Linear(11, 4)) is None
    odd = sp.ImNet(dim=3, in_features=8, out_features=4, nf=4, activation=torch.nn.Sigmoid)
    assert decoder_signature(odd) is None


# ---------------------------------------------------------------------------------------------
# numeric host logic with the torch stand-in backend (equation compiler + residual routes)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", list(RB2_CASES))
def test_pde_layer_jet_route_matches_reference_golden(name, cpu_backend):
    c = load_case(name)
    model = build_model(c, 4)
    layer = sp.get_rb2_pde_layer(**RB2_CASES[name])
    grid, q = torch.tensor(c["grid"]), torch.tensor(c["q"])
    xmin, xmax = bounds(c)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, xmin, xmax))
    with torch.no_grad():
        y, res = layer(q)
    assert rel_linf(y.numpy(), c["y_f64"]) < 1e-5
    for k, v in res.items():
        assert tuple(v.shape) == tuple(q.shape[:2]) + (1,)
        gate = max(1e-5, 2 * rel_linf(c[f"res_{k}_f32"], c[f"res_{k}_f64"]))
        assert rel_linf(v.numpy(), c[f"res_{k}_f64"]) < gate, k


@pytest.mark.parametrize("name,o", [("diffusion_leakyrelu", 2), ("ns3d_swish", 4), ("generic_d1_softplus", 2),
                                    ("generic_d2_softplus", 3), ("generic_d4_softplus", 3)])
def test_custom_equations_jet_route(name, o, cpu_backend):
    c = load_case(name)
    model = build_model(c, o)
    in_vars, out_vars, eqs = custom_equations(name, int(c["dim"]), o)
    layer = sp.PDELayer(", ".join(in_vars), ", ".join(out_vars))
    for k, (s, _) in eqs.items():
        layer.add_equation(s, k)
    grid, q = torch.tensor(c["grid"]), torch.tensor(c["q"])
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
    with torch.no_grad():
        y, res = layer(q)
    assert rel_linf(y.numpy(), c["y_f64"]) < 1e-5
    for k, v in res.items():
        assert rel_linf(v.numpy(), c[f"res_{k}_f64"]) < max(1e-5, 2 * rel_linf(c[f"res_{k}_f32"], c[f"res_{k}_f64"])), k


def test_training_gradients_match_autograd_port(cpu_backend):
    """loss.backward() through the fused Function == the reference algorithm's autograd (ref_port)."""
    from oracle import jet_oracle as jo
    from oracle import ref_port as rp

    c = load_case("rb2_softplus")
    model = build_model(c, 4)
    grid = torch.tensor(c["grid"], requires_grad=True)
    q = torch.tensor(c["q"])
    layer = sp.get_rb2_pde_layer(**RB2_CASES["rb2_softplus"])
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, 0., 1.))
    y, res = layer(q)
    loss = y.abs().mean() + 0.0125 * torch.stack(list(res.values())).abs().mean()
    loss.backward()

    port = rp.SkipMLP(c["Ws"], c["bs"], "softplus")
    grid2 = torch.tensor(c["grid"], requires_grad=True)
    iv, ov, eqs = jo.rb2_equations(**RB2_CASES["rb2_softplus"])
    y2, res2 = rp.values_and_residuals(port, grid2, q, 0., 1., iv, ov, rp.compile_equations(eqs))
    loss2 = y2.abs().mean() + 0.0125 * torch.stack(list(res2.values())).abs().mean()
    loss2.backward()
    assert abs(loss.item() - loss2.item()) < 1e-6 * abs(loss2.item())
    assert rel_linf(grid.grad.numpy(), grid2.grad.numpy()) < 1e-4
    for i in range(6):
        assert rel_linf(model.fc[i].weight.grad.numpy(), port.layers[i].weight.grad.numpy()) < 1e-4, i
        assert rel_linf(model.fc[i].bias.grad.numpy(), port.layers[i].bias.grad.numpy()) < 1e-4, i


def test_post_processed_forward_method_stays_twice_differentiable(cpu_backend):
    """A forward method that post-processes the fused output (here: 2*y - y) takes the autograd route of the
    reference; dif(dif(y, x), x) and loss.backward() must then work exactly as with the reference modules
    (the custom backward returned graph-less first derivatives)."""
    from oracle import jet_oracle as jo
    from oracle import ref_port as rp

    c = load_case("rb2_softplus")
    model = build_model(c, 4)
    grid = torch.tensor(c["grid"], requires_grad=True)
    q = torch.tensor(c["q"][:, :96])
    layer = sp.get_rb2_pde_layer(**RB2_CASES["rb2_softplus"])

    def fwd(pts):
        out = sp.query_local_implicit_grid(model, grid, pts, 0., 1.)
        return 2.0 * out - out

    layer.update_forward_method(fwd)
    y, res = layer(q)
    assert all(v.requires_grad for v in res.values())
    loss = y.abs().mean() + 0.0125 * torch.stack(list(res.values())).abs().mean()
    loss.backward()

    port = rp.SkipMLP(c["Ws"], c["bs"], "softplus")
    grid2 = torch.tensor(c["grid"], requires_grad=True)
    iv, ov, eqs = jo.rb2_equations(**RB2_CASES["rb2_softplus"])
    y2, res2 = rp.values_and_residuals(port, grid2, q, 0., 1., iv, ov, rp.compile_equations(eqs))
    loss2 = y2.abs().mean() + 0.0125 * torch.stack(list(res2.values())).abs().mean()
    loss2.backward()
    for k in res:
        assert rel_linf(res[k].detach().numpy(), res2[k].detach().numpy()) < 1e-5, k
    assert rel_linf(grid.grad.numpy(), grid2.grad.numpy()) < 1e-4
    for i in range(6):
        assert rel_linf(model.fc[i].weight.grad.numpy(), port.layers[i].weight.grad.numpy()) < 1e-4, i


@pytest.mark.parametrize("name", GRAD_CASES)
def test_torch_jet_checker_gradients_match_reference_golden(name, cpu_backend):
    """Pins the CHECKER of the GPU backward tests: loss.backward() through the jet route with the torch-op jet
    evaluator (float32 on CPU) against the real reference's float64 gradients (tests/golden/grads_*.npz)."""
    c, g = load_case(name), load_grads(name)
    o = c["Ws"][5].shape[0]
    model = build_model(c, o)
    grid = torch.tensor(c["grid"], requires_grad=True)
    q = torch.tensor(c["q"])
    xmin, xmax = bounds(c)
    layer = pde_layer_for(sp, name, c)
    layer.update_forward_method(lambda pts: sp.query_local_implicit_grid(model, grid, pts, xmin, xmax))
    y, res = layer(q, return_residue=True)
    loss = y.abs().mean() + 0.0125 * torch.stack(list(res.values())).abs().mean()
    loss.backward()
    assert abs(loss.item() - float(g["loss"])) < 1e-5 * abs(float(g["loss"]))
    assert rel_linf(grid.grad.numpy(), g["g_grid"]) < 2e-4
    for i in range(6):
        assert rel_linf(model.fc[i].weight.grad.numpy(), g[f"g_W{i}"]) < 2e-4, i
        assert rel_linf(model.fc[i].bias.grad.numpy(), g[f"g_b{i}"]) < 2e-4, i
    if "g_beta" in g:   # learnable Swish beta (reference src/nonlinearities.py:5-13)
        assert rel_linf(model.activ.beta.grad.numpy().reshape(1), g["g_beta"]) < 2e-4


def _run_postfix(words, consts, q, y, jets, gres=None):
    """Reference interpreter of the residual / adjoint programs (mirrors residual_kernel, numpy float64)."""
    outs, st = [], []
    it = iter(words)
    for op in it:
        arg = next(it)
        if op == 0: st.append(np.full(y.shape[:-1], consts[arg]))
        elif op == 1: st.append(q[..., arg])
        elif op == 2: st.append(y[..., arg])
        elif op == 3: st.append(jets[arg // y.shape[-1]][..., arg % y.shape[-1]])
        elif op == 4: b = st.pop(); st[-1] = st[-1] + b
        elif op == 5: b = st.pop(); st[-1] = st[-1] * b
        elif op == 6: st[-1] = -st[-1]
        elif op == 7: st[-1] = st[-1] ** arg
        elif op == 9: st.append(gres[arg])
        else:
            outs.append(st.pop())
            assert not st
    return outs


@pytest.mark.parametrize("kind", ["rb2", "rb2_normalised", "ns3d", "generic_d2"])
def test_adjoint_programs_match_autograd_of_the_equations(kind):
    """The symbolically differentiated adjoint programs (stpde_residuals_backward) against torch.autograd through the
    lambdified equations, on random (q, y, jets, gres)."""
    if kind == "rb2":
        layer = sp.get_rb2_pde_layer(t_crop=2., z_crop=1., x_crop=1., use_continuity=True)
    elif kind == "rb2_normalised":
        layer = sp.get_rb2_pde_layer(**RB2_CASES["rb2_paper_softplus"])
    else:
        name = "ns3d_swish" if kind == "ns3d" else "generic_d2_softplus"
        in_vars, out_vars, eqs = custom_equations(name, 2, 3)
        layer = sp.PDELayer(in_vars=", ".join(in_vars), out_vars=", ".join(out_vars))
        for eq_name, (string, subs) in eqs.items():
            layer.add_equation(string, eq_name, subs_dict=subs)
    spec, program = layer._binding()
    assert program is not None and layer._adjoint is not None
    rng = np.random.default_rng(0)
    b, p, d, o = 2, 17, layer.n_in, layer.n_out
    q = rng.normal(size=(b, p, d)); y = rng.normal(size=(b, p, o)); jets_ = rng.normal(size=(spec.n_jet, b, p, o))
    names = list(layer.eqns_raw.keys())
    gres = rng.normal(size=(len(names), b, p))
    res = _run_postfix(*program, q, y, jets_)
    qt = torch.tensor(q); yt = torch.tensor(y, requires_grad=True); jt = torch.tensor(jets_, requires_grad=True)
    total = 0.
    for e, name in enumerate(names):
        ce = layer.eqns_jet[name]
        args = []
        for s_ in ce.arg_symbols:
            if s_ in ce.jet_symbols:
                oi, multi = ce.jet_symbols[s_]
                args.append(jt[spec.plane(multi)][..., oi])
            elif s_ in layer.in_vars:
                args.append(qt[..., layer.in_vars.index(s_)])
            else:
                args.append(yt[..., layer.out_vars.index(s_)])
        val = ce.torch_fn(*args)
        assert np.allclose(res[e], val.detach().numpy(), rtol=1e-9, atol=1e-9)
        total = total + (val * torch.tensor(gres[e])).sum()
    gy, gj = torch.autograd.grad(total, [yt, jt], allow_unused=True)
    adj = _run_postfix(*layer._adjoint, q, y, jets_, gres)
    assert len(adj) == o * (1 + spec.n_jet)
    for i in range(o):
        assert np.allclose(adj[i], gy[..., i].numpy(), rtol=1e-9, atol=1e-9)
    for pl in range(spec.n_jet):
        for i in range(o):
            ref = gj[pl][..., i].numpy() if gj is not None else 0.
            assert np.allclose(adj[o + pl * o + i], ref, rtol=1e-9, atol=1e-9)


def test_backward_precision_option_and_route_selection():
    from space_time_pde_b200.equations import JetSpec
    with pytest.raises(ValueError):
        jets.set_backward_precision("bf16")
    jets.set_backward_precision("fp16")
    assert jets.BACKWARD_PRECISION == "fp16"
    jets.set_backward_precision("same")
    q_cpu = torch.zeros(1, 4, 3)
    spec = JetSpec((0, 1, 2), ((1, 1), (2, 2)))
    # the fused reverse sweep needs CUDA points, >= 3 linear layers and no gradient w.r.t. the query points
    assert not jets.fused_backward_supported(q_cpu, spec, 6, False, False)
    os.environ["STPDE_BACKWARD"] = "torch"
    try:
        assert not jets.fused_backward_supported(q_cpu, spec, 6, False, False)
    finally:
        os.environ.pop("STPDE_BACKWARD")


def test_non_polynomial_equations_keep_the_torch_route():
    """sin() cannot be expressed as a postfix program: no residual program, no adjoint program, lambdified torch arithmetic."""
    layer = sp.PDELayer(in_vars="x, t", out_vars="u")
    layer.add_equation("dif(u, t) + sin(u) * dif(u, x)", "burgers_like")
    spec, program = layer._binding()
    assert spec is not None and program is None and layer._adjoint is None


def test_deferred_checks_scope_and_late_report():
    """deferred_checks(): inside the block the calls' status words are not waited for (jets._async_mode()); leaving the
    outermost block reads every pending word and raises for the first error; an exception inside the block is not masked."""
    from space_time_pde_b200 import jets

    class Ev:                                             # stand-in for a finished CUDA event
        def synchronize(self):
            pass

        def query(self):
            return True

    assert not jets._async_mode()
    with sp.deferred_checks():
        assert jets._async_mode()
        with sp.deferred_checks():
            assert jets._async_mode()
        assert jets._async_mode()                         # still inside the outer block: nothing was read yet
    assert not jets._async_mode()
    with pytest.raises(IndexError, match="reported late"):
        with sp.deferred_checks():
            jets._pending.append((Ev(), torch.tensor([1], dtype=torch.int32), "forward"))
    assert not jets._pending and not jets._async_mode()
    with pytest.raises(ZeroDivisionError):
        with sp.deferred_checks():
            1 / 0
    assert not jets._async_mode()


def test_bench_plane_byte_accounting():
    """bench.plane_bytes_per_point: the HBM roofline of the nf = 32 legs counts every operand plane written once and read
    once (hand count for ImNet nf = 32, K = 6: 46 080 B per (point, corner) row in the parity mode; a training step in
    the single-pass mode 66 816 B, its pre-activation planes being fp16 there) and doubles with the hi + lo planes."""
    import bench
    assert bench.imnet_widths(32) == [512, 256, 128, 64, 32]
    assert bench.plane_bytes_per_point(32, 3, 6, "fp16x3") == 8 * 46080
    assert bench.plane_bytes_per_point(32, 3, 6, "fp16") == 8 * 23040
    assert bench.plane_bytes_per_point(32, 3, 6, "fp16", training=True) == 8 * 66816
    r = bench.hbm_roofline_of(1.0e7, 8 * 46080, {"hbm": 6551.0})
    assert r["bound"] == "hbm" and abs(r["frac"] - 1.0e7 * 8 * 46080 / 1e9 / 6551.0) < 1e-12
