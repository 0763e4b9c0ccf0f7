"""Procedural dataset with KNOWN factors of variation (dSprites-like: `lat_sizes`, `lat_names`, images ordered with
the last factor fastest, utils/datasets.py:148-160) for the MIG / AAM metric tests.  Shared by the golden generator
(reference on CPU) and the GPU tests; pure index arithmetic + one seeded torch.rand, so both sides build identical
tensors."""
import numpy as np
import torch


class FactorRectangles(torch.utils.data.Dataset):
    lat_names = ('posX', 'posY', 'width', 'intensity')

    def __init__(self, k=10, size=32, seed=11):
        self.lat_sizes = np.array([k, k, k, k])
        n = k ** 4
        idx = torch.arange(n)
        f3 = idx % k
        f2 = (idx // k) % k
        f1 = (idx // (k * k)) % k
        f0 = idx // (k * k * k)
        ar = torch.arange(size).view(1, size)
        x0, y0, w = (2 + 2 * f0).view(-1, 1), (2 + 2 * f1).view(-1, 1), (3 + f2).view(-1, 1)
        inten = (0.3 + 0.07 * f3.float()).view(-1, 1, 1)
        col = ((ar >= x0) & (ar < x0 + w)).float().view(n, 1, size)
        row = ((ar >= y0) & (ar < y0 + 5)).float().view(n, size, 1)
        g = torch.Generator().manual_seed(seed)
        self.imgs = (row * col * inten + 0.02 * torch.rand(n, size, size, generator=g)).clamp_(0, 1).unsqueeze(1)

    def __len__(self):
        return self.imgs.size(0)

    def __getitem__(self, i):
        return self.imgs[i], 0


def loader(ds, batch_size=500):
    return torch.utils.data.DataLoader(ds, batch_size=batch_size, shuffle=False)
