"""MIG / AAM metrics on the CUDA path (reference disvae/evaluate.py:119-317): the marginal-entropy
kernel against the reference's recorded outputs and against the oracle at scale, and disvae.Evaluator(is_metrics=True)
end to end against the reference Evaluator's metrics.log / metric_helpers.pth."""
import json
import logging
import os
import sys

import pytest
import torch

from oracle import disvae_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)


def _evaluator(model=None, save_dir="."):
    import disvae
    if model is None:
        model = disvae.init_specific_model("Burgess", (1, 32, 32), 10)
    return disvae.Evaluator(model, None, device=torch.device(DEV), logger=logging.getLogger("t"), save_dir=save_dir,
                            is_progress_bar=False)


@pytest.mark.parametrize("key", ["n3000_d10_s500", "n1200_d6_s1200"])
def test_latent_entropy_kernel_matches_reference_golden(golden, key, tmp_path):
    g = golden("metrics.pt")[key]
    ev = _evaluator(save_dir=str(tmp_path))
    ev._perm_queue = [g["perm"]]
    H = ev._estimate_latent_entropies(g["samples"].to(DEV), (g["mean"].to(DEV), g["logvar"].to(DEV)), n_samples=g["s"])
    assert torch.allclose(H.cpu(), g["H"], rtol=1e-4, atol=1e-5), (H.cpu(), g["H"])
    # interleaved (mean, logvar) views like Evaluator._compute_q_zCx produces (q_zCx.unbind(-1)): strided reads
    q = torch.stack([g["mean"], g["logvar"]], dim=-1).to(DEV)
    ev._perm_queue = [g["perm"]]
    H2 = ev._estimate_latent_entropies(g["samples"].to(DEV), q.unbind(-1), n_samples=g["s"])
    assert torch.equal(H, H2)


def test_latent_entropy_kernel_at_scale_against_oracle():
    """dSprites-sized call: N = 737280 posteriors, D = 10, S = 10000 samples (7.4e10 log-density evaluations; the
    reference materialises them 10 samples at a time).  Checked through the per-sample log q(z) output against the oracle
    on a subset of samples, plus determinism and the mean."""
    from disvae import _native as N
    torch.manual_seed(3)
    n, d, s = 737280, 10, 10000
    mean = torch.randn(n, d)
    logvar = torch.randn(n, d) * 0.7 - 1.5
    zs = torch.randn(d, s) * 1.3
    zs[:, 5] = 40.0                                             # a sample far from every posterior (no underflow to -inf)
    md, ld, zd = mean.to(DEV), logvar.to(DEV), zs.to(DEV)
    L = N.lib()
    ws = torch.empty((L.dv_latent_entropy_workspace_bytes(n, d, s) + 3) // 4, device=DEV)
    H, logq = torch.empty(d, device=DEV), torch.empty(d, s, device=DEV)
    N.call("dv_latent_entropy", N.ptr(zd), N.ptr(md), N.ptr(ld), 1, d, n, d, s, N.ptr(H), N.ptr(logq), N.ptr(ws), N.stream())
    torch.cuda.synchronize()
    assert torch.isfinite(logq).all()
    pick = torch.tensor([0, 1, 5, 77, 4095, 9999])
    ref = -torch.log(torch.tensor(float(n))) + torch.logsumexp(
        O.log_density_gaussian(zs[:, pick].unsqueeze(0).double(), mean.unsqueeze(-1).double(), logvar.unsqueeze(-1).double()), dim=0)
    got = logq.cpu()[:, pick].double()
    assert ((got - ref).abs() / ref.abs().clamp_min(1.0)).max().item() < 2e-5, (got, ref)
    assert abs(H.cpu().double() + logq.cpu().double().mean(1)).max().item() < 1e-4
    H2 = torch.empty(d, device=DEV)
    N.call("dv_latent_entropy", N.ptr(zd), N.ptr(md), N.ptr(ld), 1, d, n, d, s, N.ptr(H2), None, N.ptr(ws), N.stream())
    assert torch.equal(H, H2)


def test_evaluator_metrics_match_reference_golden(golden, tmp_path):
    """disvae.Evaluator(...)(loader, is_metrics=True, is_losses=False) == the reference Evaluator on the same procedural
    4-factor dataset and the shipped VAE_mnist checkpoint: marginal and conditional entropies, MIG, AAM, files written."""
    import disvae
    from synthetic_factors import FactorRectangles, loader
    g = golden("metrics.pt")["full"]
    ds = FactorRectangles(k=g["k"])
    d = ds.imgs.double().flatten()
    assert abs(d.sum().item() - g["imgs_digest"]["sum"]) < 1e-6 * g["imgs_digest"]["abssum"]
    torch.manual_seed(1234)
    m = disvae.init_specific_model("Burgess", (1, 32, 32), 10)
    m.load_state_dict(torch.load(os.path.join(HERE, "golden", "ckpt", "VAE_mnist.pt")))
    ev = _evaluator(m.to(DEV), save_dir=str(tmp_path))
    ev._perm_queue = list(g["perms"])
    fn = disvae.Evaluator._estimate_latent_entropies
    old = fn.__defaults__
    fn.__defaults__ = (g["n_samples"],)           # the fixture's n_samples (see tests/golden/make_golden.py:gen_metrics)
    try:
        metric, losses = ev(loader(ds), is_metrics=True, is_losses=False)
    finally:
        fn.__defaults__ = old
    assert metric is None and losses is None and not ev._perm_queue          # evaluate.py:77-79,96; every draw consumed
    metrics = json.load(open(os.path.join(str(tmp_path), "metrics.log")))
    helpers = torch.load(os.path.join(str(tmp_path), "metric_helpers.pth"), weights_only=False)
    ref = g["helpers"]
    assert set(helpers) == set(ref)
    for k in ("marginal_entropies", "cond_entropies"):
        assert tuple(helpers[k].shape) == tuple(ref[k].shape)
        assert torch.allclose(helpers[k], ref[k], rtol=1e-4, atol=1e-4), (k, (helpers[k] - ref[k]).abs().max())
    for k in ("mig_k", "aam_k", "mig", "aam"):
        assert torch.allclose(helpers[k], ref[k], rtol=1e-3, atol=1e-4), k
    assert abs(metrics["MIG"] - g["metrics"]["MIG"]) < 1e-4 and abs(metrics["AAM"] - g["metrics"]["AAM"]) < 1e-4


def test_metric_formulas_match_reference_golden(golden, tmp_path):
    g = golden("metrics.pt")["formulas"]
    ev = _evaluator(save_dir=str(tmp_path))
    st = {}
    import numpy as np
    mig = ev._mutual_information_gap(g["sorted_mut_info"], np.array([10, 10, 10, 10]), storer=st)
    aam = ev._axis_aligned_metric(g["sorted_mut_info"], storer=st)
    assert torch.allclose(mig, g["mig"]) and torch.allclose(aam, g["aam"])
    assert torch.allclose(st["mig_k"], g["mig_k"]) and torch.allclose(st["aam_k"], g["aam_k"])


def test_metrics_need_known_factors(tmp_path):
    ev = _evaluator(save_dir=str(tmp_path))

    class NoFactors(list):
        pass
    loader = NoFactors()
    loader.dataset = list(range(4))               # the reference reads dataloader.__dict__["dataset"] for its message
    with pytest.raises(ValueError):
        ev.compute_metrics(loader)
