"""Pins the oracle (oracle/disvae_oracle.py) against fixtures produced by the
unmodified reference (tests/golden/make_golden.py).  CPU only."""
import os
from collections import OrderedDict

import pytest
import torch

from oracle import disvae_oracle as O

SEED = 1234
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def digest_close(t, dg, rtol=1e-6):
    t = t.detach().double().flatten()
    assert t.numel() == dg["n"]
    for ours, ref in ((t[:8].float(), dg["head"]), (t[-8:].float(), dg["tail"])):
        # element error is judged against the magnitude of its neighbours (fp32 sums of mixed sign)
        assert torch.allclose(ours, ref, rtol=rtol, atol=1e-7 + rtol * ref.abs().max().item())
    tol = max(rtol, 1e-12) * max(1.0, dg["abssum"])      # 1e-12: fp64 summation-order slack
    assert abs(t.sum().item() - dg["sum"]) <= tol
    assert abs(t.abs().sum().item() - dg["abssum"]) <= tol


@pytest.mark.parametrize("img_size,z", [((1, 32, 32), 10), ((1, 64, 64), 10), ((3, 64, 64), 10), ((3, 64, 64), 64)])
def test_seeded_vae_init_matches_reference(golden, img_size, z):
    g = golden("init.pt")["vae_%dx%dx%d_z%d" % (img_size + (z,))]
    torch.manual_seed(SEED)
    p = O.init_vae_params(img_size, z)
    assert list(p.keys()) == g["keys"]
    for k, v in p.items():
        assert tuple(v.shape) == g["shapes"][k]
        digest_close(v, g["digest"][k], rtol=0)        # same RNG stream => bit-exact


@pytest.mark.parametrize("z", [10, 64])
def test_seeded_disc_init_matches_reference(golden, z):
    g = golden("init.pt")["disc_z%d" % z]
    torch.manual_seed(SEED)
    p = O.init_disc_params(z)
    assert list(p.keys()) == g["keys"]
    for k, v in p.items():
        digest_close(v, g["digest"][k], rtol=0)


def test_bad_image_size_raises():
    with pytest.raises(RuntimeError):
        O.init_vae_params((1, 28, 28), 10)


def _params_for(case):
    torch.manual_seed(SEED)
    p = O.init_vae_params(case["img_size"], case["latent_dim"])
    if case["ckpt"] is not None:
        sd = torch.load(os.path.join(GOLDEN, "ckpt", case["ckpt"] + ".pt"))
        p = OrderedDict((k, sd[k]) for k in p.keys())
    return p


@pytest.mark.parametrize("name", ["c1_1x32x32", "c2_1x64x64", "c3_3x64x64", "c5_3x64x64_z64",
                                  "ckpt_btcvae_dsprites", "ckpt_VAE_mnist"])
def test_forward_backward_matches_reference(golden, name):
    c = golden("forward.pt")[name]
    p = O.make_leaf_params(_params_for(c))
    mu, logvar = O.encoder_forward(p, c["x"])
    z = O.reparameterize(mu, logvar, c["eps"])
    recon = O.decoder_forward(p, z)
    assert torch.allclose(mu, c["mu"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(logvar, c["logvar"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(z, c["z"], rtol=1e-5, atol=1e-6)
    assert torch.allclose(recon, c["recon"], rtol=1e-5, atol=1e-6)
    wr = torch.linspace(0.5, 1.5, recon.numel()).view_as(recon)
    probe = (recon * wr).sum() + (mu * 0.3).sum() - (logvar * 0.2).sum()
    assert abs(probe.item() - c["probe"]) <= 1e-5 * abs(c["probe"])
    probe.backward()
    for k, v in p.items():
        digest_close(v.grad, c["grad_digest"][k], rtol=2e-5)


@pytest.mark.parametrize("loss_name", ["VAE", "betaH", "betaB", "btcvae"])
@pytest.mark.parametrize("rec_dist", ["bernoulli", "laplace", "gaussian"])
@pytest.mark.parametrize("anneal,n_calls", [(0, 1), (100, 3)])
def test_losses_match_reference(golden, loss_name, rec_dist, anneal, n_calls):
    G = golden("losses.pt")
    i = G["inputs"]
    g = G["%s_%s_a%d" % (loss_name, rec_dist, anneal)]
    assert g["n_train_steps"] == n_calls
    recon = i["recon"].clone().requires_grad_(True)
    mu = i["mu"].clone().requires_grad_(True)
    lv = i["logvar"].clone().requires_grad_(True)
    z = mu + torch.exp(0.5 * lv) * i["eps"]
    step = n_calls
    if loss_name in ("VAE", "betaH"):
        loss, logs = O.loss_betaH(i["data"], recon, mu, lv, 1 if loss_name == "VAE" else 4, rec_dist, step, anneal)
    elif loss_name == "betaB":
        loss, logs = O.loss_betaB(i["data"], recon, mu, lv, 0, 25, 100, rec_dist, step, anneal)
    else:
        loss, logs = O.loss_btcvae(i["data"], recon, mu, lv, z, 737280, 1, 6, 1, rec_dist, step, anneal)
    assert abs(loss.item() - g["loss"]) <= 1e-6 * abs(g["loss"])
    gr = torch.autograd.grad(loss, [recon, mu, lv])
    digest_close(gr[0], g["g_recon"], rtol=1e-5)
    assert torch.allclose(gr[1], g["g_mu"], rtol=1e-5, atol=1e-7)
    assert torch.allclose(gr[2], g["g_logvar"], rtol=1e-5, atol=1e-7)
    # storer of the eval-mode call (always records, anneal = 1)
    with torch.no_grad():
        ze = i["mu"] + torch.exp(0.5 * i["logvar"]) * i["eps"]
        if loss_name in ("VAE", "betaH"):
            le, logs = O.loss_betaH(i["data"], i["recon"], i["mu"], i["logvar"], 1 if loss_name == "VAE" else 4,
                                    rec_dist, step, anneal, is_train=False)
        elif loss_name == "betaB":
            le, logs = O.loss_betaB(i["data"], i["recon"], i["mu"], i["logvar"], 0, 25, 100, rec_dist, step,
                                    anneal, is_train=False)
        else:
            le, logs = O.loss_btcvae(i["data"], i["recon"], i["mu"], i["logvar"], ze, 737280, 1, 6, 1, rec_dist,
                                     step, anneal, is_train=False)
    assert abs(le.item() - g["loss_eval"]) <= 1e-6 * abs(g["loss_eval"])
    st = g["storer_eval"]
    for k, v in st.items():
        if k.startswith("kl_loss_"):
            ours = logs["kl_dims"][int(k.split("_")[-1])].item()
        else:
            ours = logs[k].item()
        assert abs(ours - v[0]) <= 1e-5 * max(1e-3, abs(v[0])), k


@pytest.mark.parametrize("key", ["b64_d10", "b256_d64", "b7_d3", "b2_d1"])
@pytest.mark.parametrize("mss", [1, 0])
def test_btcvae_density_matches_reference(golden, key, mss):
    G = golden("btcvae_density.pt")
    g = G["%s_mss%d" % (key, mss)]
    z = g["z"].clone().requires_grad_(True)
    mu = g["mu"].clone().requires_grad_(True)
    lv = g["logvar"].clone().requires_grad_(True)
    outs = O.btcvae_log_densities(z, mu, lv, g["n_data"], is_mss=bool(mss))
    for o, name in zip(outs, ["log_pz", "log_qz", "log_prod_qzi", "log_q_zCx"]):
        assert torch.allclose(o, g[name], rtol=1e-5, atol=1e-5), name
    probe = sum(c * o.mean() for c, o in zip(g["coef"], outs))
    gr = torch.autograd.grad(probe, [z, mu, lv])
    for a, b in zip(gr, [g["g_z"], g["g_mu"], g["g_logvar"]]):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)
    assert torch.equal(O.log_importance_weight_matrix(g["b"], g["n_data"]), G["logiw_b%d" % g["b"]])


def test_permute_dims_matches_reference(golden):
    g = golden("permute.pt")
    torch.manual_seed(SEED + 7)
    assert torch.equal(O.permute_dims(g["z"]), g["z_perm"])


@pytest.mark.parametrize("loss_name", ["VAE", "betaH", "betaB", "btcvae", "factor"])
def test_train_steps_match_reference(golden, loss_name):
    g = golden("train_steps.pt")[loss_name]
    over = g["over"]
    torch.set_num_threads(1)
    torch.manual_seed(SEED)
    p = O.make_leaf_params(O.init_vae_params(g["img_size"], 10))
    opt = O.make_adam(p, over["lr"])
    cfg = dict(rec_dist="bernoulli", reg_anneal=0, betaH_B=4, betaB_initC=0, betaB_finC=25, betaB_G=100,
               factor_G=6, btcvae_A=1, btcvae_B=6, btcvae_G=1, n_data=g["n_data"])
    cfg.update({k: v for k, v in over.items() if k != "lr"})
    if loss_name == "factor":
        dp = O.make_leaf_params(O.init_disc_params(10))
        opt_d = O.make_adam(dp, over["lr_disc"], betas=(0.5, 0.9))
    torch.manual_seed(SEED + 4)
    xs = [torch.rand(g["batch"], *g["img_size"]) for _ in range(3)]
    torch.manual_seed(SEED + 5)
    for i, x in enumerate(xs):
        digest_close(x, g["xs_digest"][i], rtol=0)
        if loss_name == "factor":
            loss, logs, _ = O.factor_step(p, dp, opt, opt_d, x, cfg, step=i + 1)
        else:
            loss, logs, _ = O.train_step(p, opt, x, loss_name, cfg, step=i + 1)
        ref = g["steps"][i]
        assert abs(loss.item() - ref["loss"]) <= 2e-5 * abs(ref["loss"]), (i, loss.item(), ref["loss"])
        for k, v in ref["storer"].items():     # only step 1 records (losses.py:109)
            ours = logs["kl_dims"][int(k.split("_")[-1])].item() if k.startswith("kl_loss_") else logs[k].item()
            assert abs(ours - v[0]) <= 2e-5 * max(1e-3, abs(v[0])), k
    for k, v in p.items():
        digest_close(v, g["params"][k], rtol=2e-5)
    for k, v in p.items():                        # Adam moments after 3 steps (CPU vs CPU: same trajectory)
        digest_close(opt.state[v]["exp_avg"], g["opt_state"][k]["exp_avg"], rtol=2e-5)
        digest_close(opt.state[v]["exp_avg_sq"], g["opt_state"][k]["exp_avg_sq"], rtol=2e-5)
        assert float(opt.state[v]["step"]) == g["opt_state"][k]["step"] == 3.0
    if loss_name == "factor":
        for k, v in dp.items():
            digest_close(v, g["disc_params"][k], rtol=2e-5)
            digest_close(opt_d.state[v]["exp_avg"], g["disc_opt_state"][k]["exp_avg"], rtol=2e-5)
            digest_close(opt_d.state[v]["exp_avg_sq"], g["disc_opt_state"][k]["exp_avg_sq"], rtol=2e-5)


@pytest.mark.parametrize("key", ["n3000_d10_s500", "n1200_d6_s1200"])
def test_latent_entropy_estimator_matches_reference(golden, key):
    """oracle.estimate_latent_entropies == reference Evaluator._estimate_latent_entropies (evaluate.py:233-297) on the
    recorded index draw, including the reshape-not-transpose of the selected samples."""
    g = golden("metrics.pt")[key]
    H = O.estimate_latent_entropies(g["samples"], g["mean"], g["logvar"], g["perm"])
    assert torch.allclose(H, g["H"], rtol=2e-5, atol=1e-6), (H, g["H"])


def test_metrics_tables_match_reference(golden):
    """Conditional entropies + MIG / AAM of the full fixture from the recorded marginal statistics is a GPU test (it needs
    the encoder); here: the two metric formulas (evaluate.py:163-198) incl. the 0/0 -> 0 branch."""
    g = golden("metrics.pt")["formulas"]
    smi = g["sorted_mut_info"]
    # feed a table whose sort/clamp is the identity: H_z - H_zCv == smi
    mig, aam, mig_k, aam_k = O.mig_aam(torch.zeros(10), -smi, [10, 10, 10, 10])
    assert torch.allclose(mig, g["mig"], rtol=1e-6) and torch.allclose(aam, g["aam"], rtol=1e-6)
    assert torch.allclose(mig_k, g["mig_k"], rtol=1e-6) and torch.allclose(aam_k, g["aam_k"], rtol=1e-6)
