"""API-compatibility helpers mirroring disvae/utils/math.py of the reference.

NOT on the training path: BtcvaeLoss evaluates these quantities inside the fused CUDA kernel
(csrc/dv_btcvae.cu) without materialising the B x B (x D) tensors.  They remain for callers
such as the Evaluator (out of scope) that import them by name.
"""
import math

import torch


def log_density_gaussian(x, mu, logvar):
    """reference math.py:34-51"""
    normalization = -0.5 * (math.log(2 * math.pi) + logvar)
    return normalization - 0.5 * ((x - mu) ** 2 * torch.exp(-logvar))


def matrix_log_density_gaussian(x, mu, logvar):
    """reference math.py:8-31 (materialises [B, B, D]; analysis use only)"""
    b, d = x.shape
    return log_density_gaussian(x.view(b, 1, d), mu.view(1, b, d), logvar.view(1, b, d))


def log_importance_weight_matrix(batch_size, dataset_size):
    """reference math.py:54-73 (column-structured, SURVEY.md trap T3)"""
    n, m = dataset_size, batch_size - 1
    strat = (n - m) / (n * m)
    w = torch.full((batch_size, batch_size), 1.0 / m, dtype=torch.float32)
    w.view(-1)[::m + 1] = 1.0 / n
    w.view(-1)[1::m + 1] = strat
    w[m - 1, 0] = strat
    return w.log()
