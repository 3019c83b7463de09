"""ImNet decoder with the reference's constructor, attributes and state_dict keys.

Reference: src/implicit_net.py:8-54.  The module only *holds* the parameters (so checkpoints
with keys ``fc0.*`` .. ``fc5.*`` and the aliased ``fc.0.*`` .. ``fc.5.*``, plus ``activ.beta``
for Swish, load unchanged).  On the hot path the weights are consumed directly by the fused
CUDA kernel through ``query_local_implicit_grid``; ``forward`` is the plain dense evaluation for
callers that feed pre-assembled ``[N, dim + in_features]`` rows.
"""
import torch
import torch.nn as nn


class ImNet(nn.Module):
    """Skip-MLP: widths 16nf, 8nf, 4nf, 2nf, nf, out; the input is re-appended after layers 0..3."""

    def __init__(self, dim=3, in_features=32, out_features=4, nf=32, activation=torch.nn.LeakyReLU):
        super().__init__()
        self.dim = dim
        self.in_features = in_features
        self.dimz = dim + in_features
        self.out_features = out_features
        self.nf = nf
        self.activ = activation()
        widths = [nf * 16, nf * 8, nf * 4, nf * 2, nf]
        fan_in = [self.dimz] + [w + self.dimz for w in widths[:-1]]
        layers = [nn.Linear(i, o) for i, o in zip(fan_in, widths)] + [nn.Linear(nf, out_features)]
        for idx, layer in enumerate(layers):   # same registration order / names as the reference
            setattr(self, f"fc{idx}", layer)
        self.fc = nn.ModuleList(layers)

    def forward(self, x):
        """x: [N, dim + in_features] -> [N, out_features]."""
        h = x
        last_hidden = len(self.fc) - 2
        for idx, layer in enumerate(self.fc[:-1]):
            h = self.activ(layer(h))
            if idx < last_hidden:
                h = torch.cat((h, x), dim=-1)
        return self.fc[-1](h)


def ddp_wrapper(model: nn.Module):
    """The ``DistributedDataParallel`` wrapper around ``model`` (reference train_ddp.py: ``imnet = DDP(imnet)``), or None.

    The fused path reads the inner module's parameters directly, so ``DDP.forward`` never runs and DDP's reducer is
    never armed for that step: the fused backward therefore averages the decoder gradients over the wrapper's
    process group itself (``jets.FusedJetQuery.backward``), which is what DDP would have done."""
    from torch.nn.parallel import DistributedDataParallel

    inner = model
    while isinstance(inner, nn.Module):
        if isinstance(inner, DistributedDataParallel):
            return inner
        nxt = getattr(inner, "module", None)
        if not isinstance(nxt, nn.Module):
            return None
        inner = nxt
    return None


def decoder_signature(model: nn.Module):
    """Return (layers, act_name, act_param) if ``model`` has ImNet's skip-MLP structure, else None.

    Duck-typed so that the reference's own ``implicit_net.ImNet`` (and DataParallel / DDP wrappers
    around it, reference train.py:352-355) take the fused path unchanged.
    """
    from .nonlinearities import activation_code

    inner = model
    while hasattr(inner, "module") and isinstance(getattr(inner, "module"), nn.Module):
        inner = inner.module
    fc = getattr(inner, "fc", None)
    activ = getattr(inner, "activ", None)
    if not isinstance(fc, nn.ModuleList) or activ is None or len(fc) < 2 or len(fc) > 8:
        return None
    if not all(isinstance(l, nn.Linear) and l.bias is not None for l in fc):
        return None
    dimz = fc[0].in_features
    n = len(fc)
    for i in range(1, n - 1):
        if fc[i].in_features != fc[i - 1].out_features + dimz:
            return None
    if fc[n - 1].in_features != fc[n - 2].out_features:
        return None
    if getattr(inner, "dim", None) is None or getattr(inner, "in_features", None) is None:
        return None
    if inner.dim + inner.in_features != dimz:
        return None
    name, param = activation_code(activ)
    if name is None:
        return None
    return list(fc), name, param
