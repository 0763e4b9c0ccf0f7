"""FactorVAE discriminator (reference disvae/models/discriminator.py:9-73): 6-layer MLP,
LeakyReLU(0.2), 2 logits -- one fused GEMM+bias+LeakyReLU kernel per layer."""
from torch import nn

from disvae import ops
from disvae.utils.initialization import Dense, weights_init


class Discriminator(nn.Module):
    def __init__(self, neg_slope=0.2, latent_dim=10, hidden_units=1000):
        super().__init__()
        self.neg_slope = neg_slope
        self.z_dim = latent_dim
        self.hidden_units = hidden_units
        out_units = 2
        self.lin1 = Dense(self.z_dim, hidden_units)
        self.lin2 = Dense(hidden_units, hidden_units)
        self.lin3 = Dense(hidden_units, hidden_units)
        self.lin4 = Dense(hidden_units, hidden_units)
        self.lin5 = Dense(hidden_units, hidden_units)
        self.lin6 = Dense(hidden_units, out_units)
        self.reset_parameters()

    def forward(self, z):
        params = []
        for layer in (self.lin1, self.lin2, self.lin3, self.lin4, self.lin5, self.lin6):
            params += [layer.weight, layer.bias]
        return ops.MlpFn.apply(z, self.neg_slope, *params)

    def reset_parameters(self):
        self.apply(weights_init)
