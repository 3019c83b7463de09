"""Stage the UNMODIFIED reference sources of the hot path (and of its two callers, the training and evaluation
loops) into ``oracle/_ref/`` so that they travel to the GPU box with the repository snapshot.

TEST INFRASTRUCTURE, not product: ``oracle/_ref/`` is git-ignored (the reference's sources never enter this
repository's history), it is only read by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py --impl reference``
/ the ``cpu_baseline`` leg, and everything that uses it degrades to the restatement (``oracle/ref_port.py``,
``kind: "port"``) when it is absent.  ``/root/reference`` exists only in the build container; ``__graft_entry__.build()``
calls :func:`stage` there.

Layout (mirrors the reference, so its own ``sys.path.append("../../src")`` flat imports keep working):
    oracle/_ref/src/*.py                       reference src/
    oracle/_ref/experiments/rb2d/*.py          reference experiments/rb2d/
"""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("STPDE_REFERENCE_ROOT", "/root/reference")
DEST = os.path.join(HERE, "_ref")
SUBDIRS = ("src", os.path.join("experiments", "rb2d"))


def stage(verbose: bool = False) -> bool:
    """Copy the reference's python files; returns False (and leaves things alone) when the reference is absent."""
    if not os.path.isdir(os.path.join(REF_ROOT, "src")):
        return os.path.isdir(os.path.join(DEST, "src"))
    for sub in SUBDIRS:
        src_dir, dst_dir = os.path.join(REF_ROOT, sub), os.path.join(DEST, sub)
        os.makedirs(dst_dir, exist_ok=True)
        for name in sorted(os.listdir(src_dir)):
            if name.endswith(".py"):
                shutil.copyfile(os.path.join(src_dir, name), os.path.join(dst_dir, name))
                if verbose:
                    print("staged", os.path.join(sub, name))
    return True


def available() -> bool:
    return os.path.isfile(os.path.join(DEST, "src", "pde.py"))


_STUBBED = False


def _stub_missing_modules() -> None:
    """matplotlib (train_utils / evaluation import it for colour maps) is not installed in this image: a stub module
    lets the reference files import; nothing on the measured path touches it."""
    global _STUBBED
    if _STUBBED:
        return
    import types

    import numpy as np
    if not hasattr(np, "int"):            # reference unet3d.py:191,319 use np.int (removed in numpy 2)
        np.int = int
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        for sub in ("cm", "colors", "pyplot"):
            m = types.ModuleType("matplotlib." + sub)
            setattr(mpl, sub, m)
            sys.modules["matplotlib." + sub] = m
        sys.modules["matplotlib"] = mpl
    _STUBBED = True


def import_reference(modules=("regular_nd_grid_interpolation", "local_implicit_grid", "implicit_net", "nonlinearities",
                              "pde", "physics")):
    """Import the staged reference modules under private names (``_ref_<module>``) without touching ``sys.path`` of the
    caller for good: returns {module name: module}.  The flat imports between reference files resolve against the
    staged directories while this function runs and are removed from ``sys.modules`` afterwards, so that the
    product's ``compat/`` shims of the same names are never shadowed."""
    import importlib

    if not available():
        raise ImportError("oracle/_ref is not staged (run oracle/stage_ref.py where /root/reference exists)")
    _stub_missing_modules()
    flat = ("regular_nd_grid_interpolation", "local_implicit_grid", "implicit_net", "nonlinearities", "pde", "physics",
            "unet3d", "train_utils", "dataloader_spacetime", "train", "evaluation", "torch_flow_stats",
            "torch_spec_operator", "torch_utils")
    saved = {m: sys.modules.pop(m) for m in flat if m in sys.modules}
    paths = [os.path.join(DEST, "src"), os.path.join(DEST, "experiments", "rb2d")]
    old_path = list(sys.path)
    sys.path[:0] = paths
    cwd = os.getcwd()
    out = {}
    try:
        os.chdir(paths[1])                # physics.py / train.py do sys.path.append("../../src")
        for m in modules:
            out[m] = importlib.import_module(m)
    finally:
        os.chdir(cwd)
        sys.path[:] = old_path
        for m in flat:
            mod = sys.modules.pop(m, None)
            if mod is not None:
                sys.modules["_ref_" + m] = mod
        sys.modules.update(saved)
    return out


class ReferencePipeline:
    """values + residuals through the REAL reference modules (staged copy): ``ImNet`` + ``query_local_implicit_grid`` +
    ``get_rb2_pde_layer`` / ``PDELayer`` exactly as experiments/rb2d/train.py:42-75 wires them.

    ``state_dict``: an ImNet state dict (keys fc0.* .. fc5.*, as the product's ImNet produces) so that both arms run
    the same weights."""

    def __init__(self, nf, in_features, act, state_dict, rb2_kwargs, dim=3, out_features=4, dtype=None, device="cpu"):
        import torch

        mods = import_reference()
        self.mods = mods
        self.torch = torch
        self.model = mods["implicit_net"].ImNet(dim=dim, in_features=in_features, out_features=out_features, nf=nf,
                                                activation=mods["nonlinearities"].NONLINEARITIES[act])
        self.model.load_state_dict(state_dict)
        if dtype is not None:
            self.model = self.model.to(dtype)
        self.model = self.model.to(device)
        self.layer = mods["physics"].get_rb2_pde_layer(**rb2_kwargs)

    def __call__(self, grid, q, xmin=0., xmax=1., return_residue=True):
        qlig = self.mods["local_implicit_grid"].query_local_implicit_grid
        self.layer.update_forward_method(lambda pts: qlig(self.model, grid, pts, xmin, xmax))
        return self.layer(q, return_residue=return_residue)


if __name__ == "__main__":
    ok = stage(verbose=True)
    print("oracle/_ref", "ready" if ok else "NOT staged: no reference at " + REF_ROOT)
