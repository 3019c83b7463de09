"""Activation registry with the reference's names (reference src/nonlinearities.py:5-22).

The fused kernel evaluates sigma, sigma', sigma'' of each of these in closed form
(csrc/common.cuh: act_jet); the modules below are what users construct ImNet with.
"""
import torch
import torch.nn as nn


class Swish(nn.Module):
    """x * sigmoid(beta * x) with one learnable scalar beta shared by all layers of an ImNet."""

    def __init__(self):
        super().__init__()
        self.beta = nn.Parameter(torch.tensor(1.0))

    def forward(self, x):
        return torch.sigmoid(x * self.beta) * x


NONLINEARITIES = {
    "tanh": nn.Tanh,
    "relu": nn.ReLU,
    "softplus": nn.Softplus,
    "elu": nn.ELU,
    "swish": Swish,
    "leakyrelu": nn.LeakyReLU,
}


def activation_code(module: nn.Module):
    """Map an activation module to (name, parameter tensor or None) if the fused kernel supports it."""
    name = type(module).__name__
    if name == "Tanh":
        return "tanh", None
    if name == "ReLU":
        return "relu", None
    if name == "Softplus" and float(module.beta) == 1.0 and float(module.threshold) == 20.0:
        return "softplus", None
    if name == "ELU" and float(module.alpha) == 1.0:
        return "elu", None
    if name == "Swish" and hasattr(module, "beta"):
        return "swish", module.beta
    if name == "LeakyReLU" and abs(float(module.negative_slope) - 0.01) < 1e-12:
        return "leakyrelu", None
    return None, None
