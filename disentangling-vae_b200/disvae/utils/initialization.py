"""Parameter containers and seeded initialisation.

The reference builds its layers from torch.nn.Conv2d / ConvTranspose2d / Linear
(disvae/models/encoders.py:54-67, decoders.py:53-65, discriminator.py:51-56) and then re-draws
every weight with Kaiming-uniform(relu) (disvae/utils/initialization.py:33-61).  Checkpoints
(results/*/model.pt) and `torch.manual_seed` reproducibility depend on the state_dict keys,
shapes and on the ORDER in which random numbers are consumed, so the containers below draw
exactly what the torch constructors draw (weight then bias, both U(+-1/sqrt(fan_in))) without
ever being used for compute: forward passes go through disvae.ops.
"""
import math

import torch
from torch import nn


class ParamLayer(nn.Module):
    """weight + bias holder; `fan_in` follows torch's _calculate_fan_in_and_fan_out."""

    def __init__(self, weight_shape, n_bias):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(weight_shape))
        self.bias = nn.Parameter(torch.empty(n_bias))
        fan_in = self.fan_in()
        bound = 1.0 / math.sqrt(fan_in)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)      # == kaiming_uniform_(a=sqrt(5))
            self.bias.uniform_(-bound, bound)

    def fan_in(self):
        w = self.weight
        receptive = w[0][0].numel() if w.dim() > 2 else 1
        return w.size(1) * receptive

    def forward(self, *a, **k):
        raise RuntimeError("ParamLayer only stores parameters; compute goes through disvae.ops")


class Conv4x4(ParamLayer):
    """Parameters of nn.Conv2d(c_in, c_out, 4, stride=2, padding=1): weight [c_out, c_in, 4, 4]."""

    def __init__(self, c_in, c_out):
        super().__init__((c_out, c_in, 4, 4), c_out)


class ConvT4x4(ParamLayer):
    """Parameters of nn.ConvTranspose2d(c_in, c_out, 4, stride=2, padding=1): weight [c_in, c_out, 4, 4]."""

    def __init__(self, c_in, c_out):
        super().__init__((c_in, c_out, 4, 4), c_out)


class Dense(ParamLayer):
    """Parameters of nn.Linear(n_in, n_out): weight [n_out, n_in]."""

    def __init__(self, n_in, n_out):
        super().__init__((n_out, n_in), n_out)


def linear_init(layer, activation="relu"):
    """Reference initialization.py:33-53 for the only case it is ever called with
    (activation == "relu"): Kaiming uniform, bound = sqrt(2) * sqrt(3 / fan_in)."""
    if activation != "relu":
        raise ValueError("only the reference's default activation='relu' is supported")
    bound = math.sqrt(2.0) * math.sqrt(3.0 / layer.fan_in())
    with torch.no_grad():
        layer.weight.uniform_(-bound, bound)
    return layer.weight


def weights_init(module):
    """Reference initialization.py:56-61: re-draw conv / linear weights, leave biases alone."""
    if isinstance(module, ParamLayer):
        linear_init(module)
