"""The reference's OWN training / evaluation loop (experiments/rb2d/train.py:42-183, unmodified, staged under
oracle/_ref by oracle/stage_ref.py) running on top of this package: ``compat/`` goes first on sys.path, so the loop's flat
imports (implicit_net, local_implicit_grid, pde, physics, nonlinearities) bind to the product while unet3d, train_utils
and the loop itself are the reference's files.  The first step's loss is compared with the same step evaluated by the
reference's own modules in float64 on the CPU.

Skipped when oracle/_ref is not staged (it is git-ignored; ``__graft_entry__.build()`` stages it where /root/reference
exists and it travels to the GPU box with the snapshot)."""
import argparse
import importlib
import logging
import os
import sys

import numpy as np
import pytest
import torch

from oracle import stage_ref
from space_time_pde_b200 import _torch_jets, jets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT = os.path.join(ROOT, "space_time_pde_b200", "compat")
REF_SRC = os.path.join(ROOT, "oracle", "_ref", "src")
REF_EXP = os.path.join(ROOT, "oracle", "_ref", "experiments", "rb2d")
SHARED = ("implicit_net", "local_implicit_grid", "pde", "physics", "nonlinearities", "regular_nd_grid_interpolation")
REF_ONLY = ("train", "train_utils", "unet3d", "dataloader_spacetime")

pytestmark = pytest.mark.skipif(not stage_ref.available(), reason="oracle/_ref not staged")


class Writer:
    """Stand-in for the tensorboard SummaryWriter the loop logs to: keeps the scalars."""

    def __init__(self):
        self.scalars, self.images = {}, {}

    def add_scalar(self, tag, value, global_step=None):
        self.scalars.setdefault(tag, []).append(float(value))

    def add_scalars(self, tag, values, global_step=None):
        pass

    def add_images(self, tag, images, dataformats=None, global_step=None):
        self.images[tag] = tuple(images.shape)


def import_reference_train():
    """The reference's train.py with the product's compat/ modules in front of its flat imports."""
    stage_ref._stub_missing_modules()
    for m in SHARED + REF_ONLY:
        sys.modules.pop(m, None)
    old_path, cwd = list(sys.path), os.getcwd()
    sys.path[:0] = [COMPAT, REF_SRC, REF_EXP]
    try:
        os.chdir(REF_EXP)
        train = importlib.import_module("train")
    finally:
        os.chdir(cwd)
        sys.path[:] = old_path
    assert os.path.samefile(os.path.dirname(sys.modules["physics"].__file__), COMPAT)
    assert os.path.samefile(os.path.dirname(sys.modules["local_implicit_grid"].__file__), COMPAT)
    assert os.path.samefile(os.path.dirname(train.__file__), REF_EXP)
    return train


def make_args(**kw):
    a = argparse.Namespace(reg_loss_type="l1", alpha_reg=1.0, alpha_pde=0.0125, clip_grad=1.0, log_interval=1,
                           pseudo_batch_size=1024)
    for k, v in kw.items():
        setattr(a, k, v)
    return a


def reference_first_step_loss(ref_mods, unet_state, imnet_state, batch, args, rb2, nf, lat, unet_cls, igres):
    """The same first step with the reference's own modules, float64, CPU (experiments/rb2d/train.py:56-75)."""
    unet = unet_cls(in_features=4, out_features=lat, igres=igres, nf=8, mf=32).double()
    unet.load_state_dict(unet_state)
    unet.train()
    imnet = ref_mods["implicit_net"].ImNet(dim=3, in_features=lat, out_features=4, nf=nf,
                                           activation=ref_mods["nonlinearities"].NONLINEARITIES["softplus"]).double()
    imnet.load_state_dict(imnet_state)
    layer = ref_mods["physics"].get_rb2_pde_layer(**rb2)
    input_grid, point_coord, point_value = [t.double().cpu() for t in batch]
    latent = unet(input_grid).permute(0, 2, 3, 4, 1)
    xmin, xmax = torch.zeros(3, dtype=torch.float32), torch.ones(3, dtype=torch.float32)
    qlig = ref_mods["local_implicit_grid"].query_local_implicit_grid
    layer.update_forward_method(lambda pts: qlig(imnet, latent, pts, xmin, xmax))
    pred, res = layer(point_coord, return_residue=True)
    reg = torch.nn.functional.l1_loss(pred, point_value)
    pde_t = torch.stack([d for d in res.values()], dim=0)
    pde = torch.nn.functional.l1_loss(pde_t, torch.zeros_like(pde_t))
    return float(args.alpha_reg * reg + args.alpha_pde * pde)


def run_loop(device):
    train = import_reference_train()
    ref_mods = stage_ref.import_reference()
    torch.manual_seed(0)
    nf, lat, b, p = 8, 16, 2, 96
    igres = (4, 8, 8)
    unet = train.UNet3d(in_features=4, out_features=lat, igres=igres, nf=8, mf=32)
    imnet = train.ImNet(dim=3, in_features=lat, out_features=4, nf=nf, activation=train.NONLINEARITIES["softplus"])
    assert type(imnet).__module__.startswith("space_time_pde_b200")          # the product's ImNet behind the flat name
    unet_state = {k: v.double().clone() for k, v in unet.state_dict().items()}
    imnet_state = {k: v.double().clone() for k, v in imnet.state_dict().items()}
    optimizer = torch.optim.Adam(list(unet.parameters()) + list(imnet.parameters()), lr=1e-3)
    unet = torch.nn.DataParallel(unet).to(device)
    imnet = torch.nn.DataParallel(imnet).to(device)
    rb2 = dict(mean=None, std=None, t_crop=2., z_crop=1., x_crop=1., prandtl=1., rayleigh=1e6, use_continuity=True)
    pde_layer = train.get_rb2_pde_layer(**rb2)
    gen = torch.Generator().manual_seed(1)
    batches = [(torch.randn(b, 4, *igres, generator=gen), torch.rand(b, p, 3, generator=gen),
                torch.randn(b, p, 4, generator=gen)) for _ in range(2)]
    args = make_args()
    writer = Writer()
    logger = logging.getLogger("ref_loop_test")
    before = [q.detach().clone() for q in imnet.module.parameters()]
    tot = train.train(args, unet, imnet, batches, 1, np.zeros(1, dtype=np.uint32), device, logger, writer, optimizer,
                      pde_layer)
    assert np.isfinite(tot)
    assert len(writer.scalars["train/sum_loss"]) == 2
    moved = max(float((a - q.detach()).abs().max()) for a, q in zip(before, imnet.module.parameters()))
    assert moved > 0                                                             # the optimizer stepped on the fused gradients
    want = reference_first_step_loss(ref_mods, unet_state, imnet_state, batches[0], args, rb2, nf, lat,
                                     sys.modules["_ref_unet3d"].UNet3d if "_ref_unet3d" in sys.modules else train.UNet3d, igres)
    got = writer.scalars["train/sum_loss"][0]
    assert abs(got - want) < 2e-4 * abs(want), (got, want)

    # eval(): whole slices of a high-res grid, stride-0 expanded pseudo-batches (train.py:108-183)
    train.utils.batch_colorize_scalar_tensors = lambda x, **k: torch.zeros(*x.shape, 3)   # matplotlib colour map: not the path
    hres = torch.randn(b, 4, 16, 8, 8, generator=gen)
    lres = torch.randn(b, 4, *igres, generator=gen)
    args.pseudo_batch_size = 200
    train.eval(args, unet, imnet, [(hres, lres, torch.zeros(1), torch.zeros(1))], 1, np.zeros(1, dtype=np.uint32), device,
               logger, writer, optimizer, pde_layer)
    assert writer.images["sample_0/transport_eqn_b/predicted"][:3] == (8, 8, 8)
    assert writer.images["sample_1/p/ground_truth"][:3] == (8, 8, 8)


def test_reference_train_and_eval_loop_on_the_boundary_cpu_stand_in():
    """Host logic on CPU: the CUDA kernel is replaced by the torch-op jet evaluator (tests only)."""
    jets.set_test_backend(lambda grid, q, lo, hi, Ws, bs, act, beta, spec: _torch_jets.query_jets(
        grid, q, lo, hi, list(Ws), list(bs), act, torch.tensor(beta), spec))
    try:
        run_loop(torch.device("cpu"))
    finally:
        jets.set_test_backend(None)


@pytest.mark.gpu
def test_reference_train_and_eval_loop_on_the_boundary_gpu():
    run_loop(torch.device("cuda:0"))
