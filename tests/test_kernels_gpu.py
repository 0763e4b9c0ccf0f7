"""Per-kernel parity: every C-ABI entry point against the oracle's plain-PyTorch CPU ops on the
same seeded inputs.  Tolerance: the north_star's 1e-4 relative (fp32); most kernels are far
inside it.  Run on the B200 box:  pytest -m gpu."""
import math

import pytest
import torch
import torch.nn.functional as F

from oracle import disvae_oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-4


def dev():
    return torch.device("cuda")


def rel_err(a, b):
    """max |a-b| relative to the scale of b (fp32 sums of mixed sign: judge against max |b|)."""
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def assert_close(a, b, tol=RTOL, what=""):
    assert a.shape == b.shape, (what, a.shape, b.shape)
    e = rel_err(a, b)
    assert e <= tol, "%s: rel err %.3e > %.1e" % (what, e, tol)


def nhwc(t):
    return t.permute(0, 2, 3, 1).contiguous()


def nchw(t):
    return t.permute(0, 3, 1, 2).contiguous()


@pytest.fixture(scope="module")
def ops():
    from disvae import ops as _ops
    return _ops


def relu_words(t):
    """[t > 0] of a [..., 32] tensor as one int32 word per pixel (bit c = channel c)."""
    w = ((t > 0).long() << torch.arange(32, device=t.device)).sum(-1)
    return torch.where(w >= 2 ** 31, w - 2 ** 32, w).int()


CONV_CASES = [  # (B, H(lo), CH)
    (3, 16, 1), (2, 32, 3), (5, 16, 32), (4, 8, 32), (7, 4, 32), (3, 2 * 2, 32), (2, 32, 1), (1, 16, 3),
    (170, 16, 32),      # 340 tiles of 128 pixels: several tiles per persistent CTA (pipeline phase wrap-around)
    (301, 8, 32), (1201, 4, 32),
    (40, 32, 1), (40, 32, 3),      # image-boundary layers, 320 tiles
    (100, 32, 1),                  # halo up kernel: 1100 tiles, > 5 per CTA (every shared-memory stage is reused)
    (330, 32, 1), (300, 16, 3),    # col2im up kernel: several images per persistent CTA (carry row reset between images)
]


@pytest.mark.parametrize("B,H,CH", CONV_CASES)
@pytest.mark.parametrize("act", [0, 1])
def test_conv_down_matches_conv2d(ops, B, H, CH, act):
    torch.manual_seed(B * 100 + H + CH)
    x = torch.randn(B, CH, 2 * H, 2 * H)
    w = torch.randn(32, CH, 4, 4) * 0.1
    b = torch.randn(32)
    ref = F.conv2d(x, w, b, stride=2, padding=1)
    if act:
        ref = torch.relu(ref)
    wp = ops.conv_pack(w.to(dev()), CH)
    hi = x.to(dev()) if CH < 32 else nhwc(x).to(dev())
    lo = ops.conv_down(hi, wp, b.to(dev()), None, B, H, H, CH, int(CH < 32), act)
    assert_close(nchw(lo.cpu()), ref, what="down")
    # mask epilogue (ReLU backward of the producer of `mask`)
    mask = torch.randn(B, 32, H, H)
    lo2 = ops.conv_down(hi, wp, None, nhwc(mask).to(dev()), B, H, H, CH, int(CH < 32), 0)
    ref2 = F.conv2d(x, w, None, stride=2, padding=1) * (mask > 0)
    assert_close(nchw(lo2.cpu()), ref2, what="down+mask")
    # channel sums of the output from the same launch (bias gradient of the previous ConvTranspose2d)
    lo3, cs = ops.conv_down(hi, wp, None, nhwc(mask).to(dev()), B, H, H, CH, int(CH < 32), 0, want_colsum=True)
    assert torch.equal(lo3, lo2)
    assert_close(cs.cpu(), ref2.double().sum((0, 2, 3)).float(), tol=2e-5, what="down+mask column sums")
    _, cs2 = ops.conv_down(hi, wp, None, nhwc(mask).to(dev()), B, H, H, CH, int(CH < 32), 0, want_colsum=True)
    assert torch.equal(cs, cs2)                                # fixed reduction order
    # ReLU masks as one word per pixel: produced by the forward epilogue, consumed instead of the 128-byte float rows
    lo4, bits = ops.conv_down(hi, wp, b.to(dev()), None, B, H, H, CH, int(CH < 32), act, want_bits=True)
    assert torch.equal(lo4, lo)
    assert torch.equal(bits, relu_words(lo4))
    mbits = relu_words(nhwc(mask).to(dev()))
    lo5, cs5 = ops.conv_down(hi, wp, None, nhwc(mask).to(dev()), B, H, H, CH, int(CH < 32), 0, want_colsum=True, mask_bits=mbits)
    assert torch.equal(lo5, lo2) and torch.equal(cs5, cs)


@pytest.mark.parametrize("B,H,CH", CONV_CASES)
@pytest.mark.parametrize("act", [0, 1, 2])
def test_conv_up_matches_conv_transpose2d(ops, B, H, CH, act):
    torch.manual_seed(B * 100 + H + CH + 7)
    lo = torch.randn(B, 32, H, H)
    w = torch.randn(32, CH, 4, 4) * 0.1
    b = torch.randn(CH)
    ref = F.conv_transpose2d(lo, w, b, stride=2, padding=1)
    ref = torch.relu(ref) if act == 1 else (torch.sigmoid(ref) if act == 2 else ref)
    wp = ops.conv_pack(w.to(dev()), CH)
    hi = ops.conv_up(nhwc(lo).to(dev()), wp, b.to(dev()), None, B, H, H, CH, int(CH < 32), act)
    got = hi.cpu() if CH < 32 else nchw(hi.cpu())
    assert_close(got, ref, what="up")
    if CH == 32:
        mask = torch.randn(B, 32, 2 * H, 2 * H)
        hi2 = ops.conv_up(nhwc(lo).to(dev()), wp, None, nhwc(mask).to(dev()), B, H, H, CH, 0, 0)
        ref2 = F.conv_transpose2d(lo, w, None, stride=2, padding=1) * (mask > 0)
        assert_close(nchw(hi2.cpu()), ref2, what="up+mask")
        hi3 = ops.conv_up(nhwc(lo).to(dev()), wp, None, nhwc(mask).to(dev()), B, H, H, CH, 0, 0,
                          mask_bits=relu_words(nhwc(mask).to(dev())))
        assert torch.equal(hi3, hi2)
        if act != 2:
            hi4, bits = ops.conv_up(nhwc(lo).to(dev()), wp, b.to(dev()), None, B, H, H, CH, 0, act, want_bits=True)
            assert torch.equal(hi4, hi) and torch.equal(bits, relu_words(hi4))


@pytest.mark.parametrize("B,H,CH", CONV_CASES + [(64, 16, 32), (33, 32, 3)])
def test_conv_wgrad_matches_autograd(ops, B, H, CH):
    torch.manual_seed(B * 100 + H + CH + 13)
    x = torch.randn(B, CH, 2 * H, 2 * H)
    w = torch.zeros(32, CH, 4, 4, requires_grad=True)
    b = torch.zeros(32, requires_grad=True)
    g = torch.randn(B, 32, H, H)
    (F.conv2d(x, w, b, stride=2, padding=1) * g).sum().backward()
    hi = x.to(dev()) if CH < 32 else nhwc(x).to(dev())
    dw, db = ops.conv_wgrad(nhwc(g).to(dev()), hi, B, H, H, CH, int(CH < 32), True)
    assert_close(dw.cpu(), w.grad, what="dw")
    assert_close(db.cpu(), b.grad, what="db")
    # determinism of the split-K reduction
    dw2, _ = ops.conv_wgrad(nhwc(g).to(dev()), hi, B, H, H, CH, int(CH < 32), True)
    assert torch.equal(dw, dw2)


@pytest.mark.parametrize("B,H,CH", [(5, 32, 1), (333, 32, 1), (150, 32, 3), (77, 16, 1), (200, 16, 3)])
def test_conv_up_small_fp32_grade_accuracy(ops, B, H, CH):
    """Decoder output layer (col2im tensor-core kernel): <= 4e-6 of the output scale against fp64, pre-activation
    and through the sigmoid."""
    torch.manual_seed(B + H + CH)
    lo = torch.randn(B, 32, H, H)
    w = torch.randn(32, CH, 4, 4) * 0.1
    b = torch.randn(CH)
    wp = ops.conv_pack(w.to(dev()), CH)
    ref = F.conv_transpose2d(lo.double(), w.double(), b.double(), stride=2, padding=1)
    got = ops.conv_up(nhwc(lo).to(dev()), wp, b.to(dev()), None, B, H, H, CH, 1, 0).cpu()
    assert_close(got, ref, tol=4e-6, what="up small vs fp64")
    got_s = ops.conv_up(nhwc(lo).to(dev()), wp, b.to(dev()), None, B, H, H, CH, 1, 2).cpu()
    assert_close(got_s, torch.sigmoid(ref), tol=4e-6, what="up small sigmoid vs fp64")


@pytest.mark.parametrize("B,H", [(16, 16), (9, 8), (40, 4), (200, 16), (330, 16), (700, 8)])
def test_conv32_kernels_fp32_grade_accuracy(ops, B, H):
    """The tensor-core (3xTF32) path must stay at fp32-grade accuracy, not tf32-grade: error against an
    fp64 reference <= 4e-6 of the output scale (plain fp32 lands around 5e-7, single-pass tf32 at 5e-4)."""
    torch.manual_seed(B + H)
    x = torch.randn(B, 32, 2 * H, 2 * H)
    lo = torch.randn(B, 32, H, H)
    w = torch.randn(32, 32, 4, 4) * 0.1
    wp = ops.conv_pack(w.to(dev()), 32)
    ref_d = F.conv2d(x.double(), w.double(), None, stride=2, padding=1)
    got_d = nchw(ops.conv_down(nhwc(x).to(dev()), wp, None, None, B, H, H, 32, 0, 0).cpu())
    assert_close(got_d, ref_d, tol=4e-6, what="down vs fp64")
    ref_u = F.conv_transpose2d(lo.double(), w.double(), None, stride=2, padding=1)
    got_u = nchw(ops.conv_up(nhwc(lo).to(dev()), wp, None, None, B, H, H, 32, 0, 0).cpu())
    assert_close(got_u, ref_u, tol=4e-6, what="up vs fp64")
    wz = torch.zeros(32, 32, 4, 4, dtype=torch.float64, requires_grad=True)
    (F.conv2d(x.double(), wz, None, stride=2, padding=1) * lo.double()).sum().backward()
    dw, db = ops.conv_wgrad(nhwc(lo).to(dev()), nhwc(x).to(dev()), B, H, H, 32, 0, True)
    assert_close(dw.cpu(), wz.grad, tol=4e-6, what="wgrad vs fp64")
    assert_close(db.cpu(), lo.double().sum((0, 2, 3)), tol=4e-6, what="dbias vs fp64")


def test_conv_transpose_weight_gradient_is_same_kernel(ops):
    """dW of ConvTranspose2d(32 -> CH) == wgrad(lo = its input, hi = grad of its output)."""
    torch.manual_seed(3)
    for CH in (3, 32):
        B, H = 4, 8
        lo = torch.randn(B, 32, H, H)
        w = torch.zeros(32, CH, 4, 4, requires_grad=True)
        g = torch.randn(B, CH, 2 * H, 2 * H)
        (F.conv_transpose2d(lo, w, None, stride=2, padding=1) * g).sum().backward()
        hi = g.to(dev()) if CH < 32 else nhwc(g).to(dev())
        dw, _ = ops.conv_wgrad(nhwc(lo).to(dev()), hi, B, H, H, CH, int(CH < 32), False)
        assert_close(dw.cpu(), w.grad, what="convT dw CH=%d" % CH)


def test_channel_sum_and_transpose_and_act_bwd(ops):
    torch.manual_seed(5)
    x = torch.randn(1000, 32)
    assert_close(ops.channel_sum(x.to(dev()), 1000, 32, 0, 0).cpu(), x.sum(0), tol=1e-5)
    y = torch.randn(37, 3, 64 * 64)
    assert_close(ops.channel_sum(y.to(dev()), 37, 3, 1, 64 * 64).cpu(), y.sum((0, 2)), tol=1e-5)
    t = torch.randn(9, 32, 16)                                 # [B, C, S] NCHW-flat
    got = ops.flat_transpose(t.view(9, 512).to(dev()), 9, to_nhwc=True).cpu().view(9, 16, 32)
    assert torch.equal(got, t.permute(0, 2, 1))
    back = ops.flat_transpose(got.reshape(9, 512).to(dev()), 9, to_nhwc=False).cpu().view(9, 32, 16)
    assert torch.equal(back, t)
    yy = torch.sigmoid(torch.randn(4096) * 4)
    dy = torch.randn(4096)
    assert_close(ops.act_bwd(dy.to(dev()), yy.to(dev()), 2).cpu(), dy * (1 - yy) * yy, tol=1e-6)


LIN_CASES = [(64, 256, 512), (7, 20, 256), (130, 256, 10), (33, 1000, 1000), (256, 2, 1000), (5, 128, 64), (1, 512, 256),
             (1024, 256, 512), (1000, 20, 256), (513, 256, 10)]


@pytest.mark.parametrize("M,N,K", LIN_CASES)
def test_linear_fwd_dgrad_wgrad(ops, M, N, K):
    torch.manual_seed(M + N + K)
    x = torch.randn(M, K, requires_grad=True)
    w = (torch.randn(N, K) / math.sqrt(K)).requires_grad_(True)
    b = torch.randn(N, requires_grad=True)
    for act, slope, f in [(0, 0.0, lambda t: t), (1, 0.0, torch.relu), (3, 0.2, lambda t: F.leaky_relu(t, 0.2))]:
        y = ops.linear_fwd(x.detach().to(dev()), w.detach().to(dev()), b.detach().to(dev()), act, slope)
        assert_close(y.cpu(), f(F.linear(x, w, b)), what="fwd act %d" % act)
    g = torch.randn(M, N)
    x.grad = w.grad = b.grad = None
    (F.linear(x, w, b) * g).sum().backward()
    dx = ops.linear_dgrad(g.to(dev()), w.detach().to(dev()), None, 0)
    assert_close(dx.cpu(), x.grad, what="dgrad")
    prev = torch.randn(M, K)
    dxm = ops.linear_dgrad(g.to(dev()), w.detach().to(dev()), torch.relu(prev).to(dev()), 1)
    assert_close(dxm.cpu(), x.grad * (prev > 0), what="dgrad relu mask")
    dxl = ops.linear_dgrad(g.to(dev()), w.detach().to(dev()), F.leaky_relu(prev, 0.2).to(dev()), 3, 0.2)
    assert_close(dxl.cpu(), x.grad * torch.where(prev > 0, 1.0, 0.2), what="dgrad leaky mask")
    dw, db = ops.linear_wgrad(g.to(dev()), x.detach().to(dev()))
    assert_close(dw.cpu(), w.grad, what="wgrad")
    assert_close(db.cpu(), b.grad, what="bgrad")


@pytest.mark.parametrize("M,N,K", [(1024, 256, 512), (256, 1000, 1000), (300, 1000, 12), (77, 512, 256)])
def test_linear_tensor_core_accuracy_vs_fp64(ops, M, N, K):
    """The tcgen05 linear layers use the same error-compensated 3xTF32 scheme as the convolutions: <= 4e-6 of the
    output scale against an fp64 reference (plain fp32 ~5e-7, single-pass tf32 ~5e-4)."""
    torch.manual_seed(M * 7 + N + K)
    x = torch.randn(M, K)
    w = torch.randn(N, K) / math.sqrt(K)
    b = torch.randn(N)
    g = torch.randn(M, N)
    y = ops.linear_fwd(x.to(dev()), w.to(dev()), b.to(dev()), 0)
    assert_close(y.cpu(), F.linear(x.double(), w.double(), b.double()), tol=4e-6, what="fwd vs fp64")
    dx = ops.linear_dgrad(g.to(dev()), w.to(dev()), None, 0)
    assert_close(dx.cpu(), g.double() @ w.double(), tol=4e-6, what="dgrad vs fp64")
    dw, db = ops.linear_wgrad(g.to(dev()), x.to(dev()))
    assert_close(dw.cpu(), g.double().t() @ x.double(), tol=4e-6, what="wgrad vs fp64")
    assert_close(db.cpu(), g.double().sum(0), tol=4e-6, what="dbias vs fp64")


@pytest.mark.parametrize("dist", ["bernoulli", "gaussian", "laplace"])
def test_vae_loss_kernel(ops, golden, dist):
    from disvae._native import DIST
    i = golden("losses.pt")["inputs"]
    data, recon0, mu0, lv0 = i["data"], i["recon"], i["mu"], i["logvar"]
    ml = torch.stack([mu0, lv0], dim=-1).reshape(mu0.size(0), -1)        # interleaved like the encoder output
    recon = recon0.clone().requires_grad_(True)
    mu = mu0.clone().requires_grad_(True)
    lv = lv0.clone().requires_grad_(True)
    rec = O.reconstruction_loss(data, recon, dist)
    kl, kl_dims = O.kl_normal(mu, lv)
    (1.7 * rec + 0.3 * kl).backward()
    mld = ml.to(dev()).requires_grad_(True)
    mud, lvd = mld.view(-1, mu0.size(1), 2).unbind(-1)
    rd = recon0.to(dev()).requires_grad_(True)
    out = ops.VaeLossFn.apply(rd, data.to(dev()), mud, lvd, DIST[dist])
    assert_close(out[0:1].cpu(), rec.detach().view(1), tol=2e-6, what="recon loss")
    assert_close(out[1:2].cpu(), kl.detach().view(1), tol=2e-6, what="kl")
    assert_close(out[2:].cpu(), kl_dims.detach(), tol=2e-6, what="kl dims")
    (1.7 * out[0] + 0.3 * out[1]).backward()
    assert_close(rd.grad.cpu(), recon.grad, tol=1e-5, what="d recon")
    g_ml = mld.grad.cpu().view(-1, mu0.size(1), 2)
    assert_close(g_ml[..., 0], mu.grad, tol=1e-5, what="d mu")
    assert_close(g_ml[..., 1], lv.grad, tol=1e-5, what="d logvar")


def test_laplace_zero_loss_mask(ops):
    from disvae._native import DIST
    x = torch.rand(2, 1, 32, 32, device=dev())
    r = x.clone().requires_grad_(True)
    z = torch.zeros(2, 3, device=dev())
    out = ops.VaeLossFn.apply(r, x, z, z, DIST["laplace"])
    assert out[0].item() == 0.0
    out[0].backward()
    assert torch.count_nonzero(r.grad).item() == 0


def test_reparam_fwd_bwd_and_device_noise(ops):
    torch.manual_seed(11)
    B, D = 37, 10
    ml = torch.randn(B, 2 * D)
    eps = torch.randn(B, D)
    mlc = ml.clone().requires_grad_(True)
    mu, lv = mlc.view(B, D, 2).unbind(-1)
    z = O.reparameterize(mu, lv, eps)
    (z * torch.arange(B * D).view(B, D).float()).sum().backward()
    mld = ml.to(dev()).requires_grad_(True)
    mud, lvd = mld.view(B, D, 2).unbind(-1)
    zd = ops.ReparamFn.apply(mud, lvd, eps.to(dev()), 0, None)
    assert_close(zd.cpu(), z.detach(), tol=1e-6)
    (zd * torch.arange(B * D, device=dev()).view(B, D).float()).sum().backward()
    assert_close(mld.grad.cpu(), mlc.grad, tol=1e-6)
    # device Philox noise: N(0,1) moments, reproducible per (seed, offset), offset advances
    Bn, Dn = 4096, 64
    zeros = torch.zeros(Bn, Dn, device=dev())
    off = torch.zeros(1, dtype=torch.int64, device=dev())
    e1 = ops.ReparamFn.apply(zeros, zeros, None, 1234, off)
    assert off.item() == Bn * Dn
    e2 = ops.ReparamFn.apply(zeros, zeros, None, 1234, off)
    off.zero_()
    e3 = ops.ReparamFn.apply(zeros, zeros, None, 1234, off)
    assert torch.equal(e1, e3) and not torch.equal(e1, e2)
    assert abs(e1.mean().item()) < 0.01 and abs(e1.std().item() - 1) < 0.01
    assert abs((e1 ** 3).mean().item()) < 0.03 and abs((e1 ** 4).mean().item() - 3) < 0.1


BT_CASES = ["b64_d10", "b256_d64", "b7_d3", "b2_d1"]


@pytest.mark.parametrize("key", BT_CASES)
@pytest.mark.parametrize("mss", [1, 0])
def test_btcvae_kernel_against_reference_golden(ops, golden, key, mss):
    g = golden("btcvae_density.pt")["%s_mss%d" % (key, mss)]
    B, D = g["b"], g["d"]
    z = g["z"].to(dev()).requires_grad_(True)
    ml = torch.stack([g["mu"], g["logvar"]], dim=-1).reshape(B, -1).to(dev()).requires_grad_(True)
    mu, lv = ml.view(B, D, 2).unbind(-1)
    stats = ops.btcvae_rowstats(z, mu, lv, g["n_data"], bool(mss))
    for got, name in zip(stats, ["log_pz", "log_qz", "log_prod_qzi", "log_q_zCx"]):
        assert_close(got.cpu(), g[name], tol=2e-5, what=name)
    terms = ops.BtcvaeFn.apply(z, mu, lv, g["n_data"], bool(mss))
    c = g["coef"]          # probe = c0*mean(log_pz) + c1*mean(log_qz) + c2*mean(log_prod) + c3*mean(log_qzCx)
    mi_ref = (g["log_q_zCx"] - g["log_qz"]).mean()
    tc_ref = (g["log_qz"] - g["log_prod_qzi"]).mean()
    dw_ref = (g["log_prod_qzi"] - g["log_pz"]).mean()
    assert_close(terms.cpu(), torch.stack([mi_ref, tc_ref, dw_ref]), tol=5e-5, what="terms")
    # probe in terms of (mi, tc, dw): a*mi + b*tc + e*dw with lqc coef = a = c3; lqz: -a + b = c1; lprod: -b + e = c2;
    # lpz: -e = c0  => only consistent if c0+c1+c2+c3 == 0; use autograd on the oracle instead.
    zo = g["z"].clone().requires_grad_(True)
    muo = g["mu"].clone().requires_grad_(True)
    lvo = g["logvar"].clone().requires_grad_(True)
    mi, tc, dw = O.btcvae_terms(zo, muo, lvo, g["n_data"], bool(mss))
    (1.0 * mi + 6.0 * tc - 2.5 * dw).backward()
    (1.0 * terms[0] + 6.0 * terms[1] - 2.5 * terms[2]).backward()
    assert_close(z.grad.cpu(), zo.grad, what="g_z")
    gml = ml.grad.cpu().view(B, D, 2)
    assert_close(gml[..., 0], muo.grad, what="g_mu")
    assert_close(gml[..., 1], lvo.grad, what="g_logvar")


@pytest.mark.parametrize("B,D", [(1024, 10), (256, 64), (512, 16), (129, 5), (1000, 8), (96, 20), (2048, 64)])
def test_btcvae_kernel_against_oracle_sizes(ops, B, D):
    torch.manual_seed(B + D)
    mu = torch.randn(B, D)
    lv = torch.randn(B, D) * 0.5 - 1
    z = mu + torch.exp(0.5 * lv) * torch.randn(B, D)
    n_data = 737280
    zo, muo, lvo = [t.clone().requires_grad_(True) for t in (z, mu, lv)]
    big = B * B * D > 5e7
    if big:           # oracle materialises B*B*D floats: keep the big case forward-only + properties
        with torch.no_grad():
            ref = O.btcvae_log_densities(z, mu, lv, n_data)
    else:
        ref = O.btcvae_log_densities(zo, muo, lvo, n_data)
    zd, mud, lvd = [t.to(dev()).requires_grad_(True) for t in (z, mu, lv)]
    stats = ops.btcvae_rowstats(zd, mud, lvd, n_data, True)
    for got, r, name in zip(stats, ref, ["log_pz", "log_qz", "log_prod_qzi", "log_q_zCx"]):
        assert_close(got.cpu(), r.detach(), tol=2e-5, what=name)
    terms = ops.BtcvaeFn.apply(zd, mud, lvd, n_data, True)
    (terms[0] + 6 * terms[1] + terms[2]).backward()
    if not big:
        mi = (ref[3] - ref[1]).mean(); tc = (ref[1] - ref[2]).mean(); dw = (ref[2] - ref[0]).mean()
        (mi + 6 * tc + dw).backward()
        assert_close(zd.grad.cpu(), zo.grad, what="g_z")
        assert_close(mud.grad.cpu(), muo.grad, what="g_mu")
        assert_close(lvd.grad.cpu(), lvo.grad, what="g_logvar")
    # size-independent property: run-to-run bit-exactness (fixed reduction order)
    terms2 = ops.BtcvaeFn.apply(zd, mud, lvd, n_data, True)
    assert torch.equal(terms, terms2)


def test_btcvae_extreme_variances_stay_finite(ops):
    """Trained models have tiny variances: logsumexp must not under/overflow (exact max)."""
    torch.manual_seed(0)
    B, D = 128, 10
    mu = torch.randn(B, D) * 3
    lv = torch.full((B, D), -14.0)
    lv[::7] = 3.0
    z = mu + torch.exp(0.5 * lv) * torch.randn(B, D)
    ref = O.btcvae_log_densities(z, mu, lv, 10000)
    stats = ops.btcvae_rowstats(z.to(dev()), mu.to(dev()), lv.to(dev()), 10000, True)
    for got, r in zip(stats, ref):
        assert torch.isfinite(got).all()
        assert_close(got.cpu(), r, tol=2e-5)


def test_btcvae_outlier_rows_and_tiny_variances(ops):
    """Rows far (100s of sigma) from every other column, tiny and huge variances mixed: the single-sweep
    logsumexp (bounded reference exponent) must agree with the exact-max oracle and stay finite."""
    torch.manual_seed(1)
    B, D = 192, 10
    mu = torch.randn(B, D)
    lv = torch.randn(B, D) * 0.3 - 2.0
    lv[5] = -20.0; lv[77] = 6.0; lv[130, :3] = -16.0
    z = mu + torch.exp(0.5 * lv) * torch.randn(B, D)
    z[9] += 40.0; z[100, 2] -= 25.0
    ref = O.btcvae_log_densities(z, mu, lv, 50000)
    stats = ops.btcvae_rowstats(z.to(dev()), mu.to(dev()), lv.to(dev()), 50000, True)
    for got, r, name in zip(stats, ref, ["log_pz", "log_qz", "log_prod_qzi", "log_q_zCx"]):
        assert torch.isfinite(got).all(), name
        assert_close(got.cpu(), r, tol=2e-5, what=name)


def test_permute_dims(ops, golden):
    g = golden("permute.pt")
    torch.manual_seed(1234 + 7)
    perms = torch.stack([torch.randperm(16) for _ in range(10)])
    got = ops.permute_dims(g["z"].to(dev()), perms)
    assert torch.equal(got.cpu(), g["z_perm"])
    # device-generated permutations: every column is a permutation of the input column
    B, D = 1000, 12
    z = torch.randn(B, D, device=dev())
    off = torch.zeros(1, dtype=torch.int64, device=dev())
    p1 = ops.permute_dims(z, None, 99, off)
    assert off.item() == B * D
    assert torch.equal(p1.sort(0).values, z.sort(0).values)
    assert not torch.equal(p1, z)
    p2 = ops.permute_dims(z, None, 99, off)
    assert not torch.equal(p1, p2)
    # uniformity smoke: mean displacement of a uniform random permutation ~ B/3
    idx = torch.arange(B, device=dev(), dtype=torch.float32).unsqueeze(1).repeat(1, D)
    pi = ops.permute_dims(idx, None, 5, off)
    disp = (pi - idx).abs().mean().item()
    assert abs(disp - B / 3) < 0.05 * B


def test_factor_heads(ops):
    torch.manual_seed(2)
    h = 130
    dz = torch.randn(h, 2, requires_grad=True)
    dp = torch.randn(h, 2, requires_grad=True)
    tc = (dz[:, 0] - dz[:, 1]).mean()
    ones = torch.ones(h, dtype=torch.long)
    ce = 0.5 * (F.cross_entropy(dz, torch.zeros_like(ones)) + F.cross_entropy(dp, ones))
    (2 * tc + 3 * ce).backward()
    dzd = dz.detach().to(dev()).requires_grad_(True)
    dpd = dp.detach().to(dev()).requires_grad_(True)
    tcd = ops.FactorTcFn.apply(dzd)
    ced = ops.FactorCeFn.apply(dzd, dpd)
    assert abs(tcd.item() - tc.item()) < 1e-6 and abs(ced.item() - ce.item()) < 1e-6
    (2 * tcd + 3 * ced).backward()
    assert_close(dzd.grad.cpu(), dz.grad, tol=1e-5)
    assert_close(dpd.grad.cpu(), dp.grad, tol=1e-5)


def test_adam_step_matches_torch(ops):
    torch.manual_seed(4)
    n = 100003
    p = torch.randn(n)
    po = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([po], lr=5e-4)
    pd = p.to(dev())
    m = torch.zeros(n, device=dev()); v = torch.zeros(n, device=dev()); step = torch.zeros(1, device=dev())
    for i in range(3):
        g = torch.randn(n)
        po.grad = g.clone()
        opt.step()
        ops.adam_step(pd, g.to(dev()), m, v, step, 5e-4, (0.9, 0.999), 1e-8)
    assert step.item() == 3
    assert_close(pd.cpu(), po.detach(), tol=1e-6)


def test_fused_adam_multi_matches_torch_adam():
    """disvae.fused.FusedAdam (dv_adam_multi, one launch for all tensors) == torch.optim.Adam, incl. shared state."""
    from disvae.fused import FusedAdam
    torch.manual_seed(6)
    shapes = [(32, 1, 4, 4), (32,), (256, 512), (20, 256), (1000, 1000), (3,)]
    ps_ref = [torch.randn(s, requires_grad=True) for s in shapes]
    ps = [p.detach().clone().to(dev()).requires_grad_(True) for p in ps_ref]
    opt_ref = torch.optim.Adam(ps_ref, lr=5e-4)
    opt = torch.optim.Adam(ps, lr=5e-4)
    assert FusedAdam.supports(opt) and not FusedAdam.supports(torch.optim.Adam(ps, lr=1e-3, weight_decay=0.1))
    fused = FusedAdam(opt)
    for _ in range(4):
        for a, b in zip(ps_ref, ps):
            g = torch.randn(a.shape)
            a.grad = g.clone(); b.grad = g.to(dev())
        opt_ref.step(); fused.step()
    for a, b in zip(ps_ref, ps):
        assert_close(b.detach().cpu(), a.detach(), tol=1e-6)
        assert_close(opt.state[b]["exp_avg_sq"].cpu(), opt_ref.state[a]["exp_avg_sq"], tol=1e-6)
    fused.flush_state()
    assert float(opt.state[ps[0]]["step"]) == 4.0


@pytest.mark.parametrize("B,D,world", [(256, 64, 2), (96, 10, 3), (2048, 64, 8), (130, 5, 2)])
def test_btcvae_row_windows_compose_to_the_full_batch(ops, B, D, world):
    """SURVEY.md 8f-1 kernels: the estimator over row windows of one (all-gathered) batch.  Window r evaluates rows
    [r*b, (r+1)*b) against ALL B columns; rowstats of the windows == the oracle's full-batch values; the mean of the
    windows' terms == the full-batch terms; g_z windows stack to the full g_z and the column-side partial sums add up to
    the full g_mu / g_logvar (what the reduce-scatter over ranks computes)."""
    from disvae import _native as N
    torch.manual_seed(B + D + world)
    b = B // world
    B = b * world
    mu = torch.randn(B, D)
    lv = torch.randn(B, D) * 0.5 - 1
    z = mu + torch.exp(0.5 * lv) * torch.randn(B, D)
    n_data = 202599
    big = B * B * D > 5e7
    zo, muo, lvo = [t.clone().requires_grad_(not big) for t in (z, mu, lv)]
    with torch.set_grad_enabled(not big):
        ref = O.btcvae_log_densities(zo, muo, lvo, n_data)
    coef = (1.0, 6.0, -2.5)
    if not big:
        mi, tc, dw = (ref[3] - ref[1]).mean(), (ref[1] - ref[2]).mean(), (ref[2] - ref[0]).mean()
        (coef[0] * mi + coef[1] * tc + coef[2] * dw).backward()
    zd, mud, lvd = z.to(dev()), mu.to(dev()), lv.to(dev())
    nbytes = N.lib().dv_btcvae_workspace_bytes(B, D)
    g_terms = torch.tensor(coef, device=dev())
    terms_mean = torch.zeros(3)
    g_z, g_mu, g_lv = torch.zeros(B, D), torch.zeros(B, D), torch.zeros(B, D)
    for r in range(world):
        ws = torch.zeros((nbytes + 3) // 4, device=dev())
        rowstats = torch.full((4 + D, B), float("nan"), device=dev())
        terms = torch.empty(3, device=dev())
        N.call("dv_btcvae_fwd_rows", N.ptr(zd), N.ptr(mud), N.ptr(lvd), 1, D, B, D, r * b, b, n_data, 1, N.ptr(rowstats),
               N.ptr(terms), N.ptr(ws), N.stream())
        for got, rf, name in zip(rowstats[:4], ref, ["log_pz", "log_qz", "log_prod_qzi", "log_q_zCx"]):
            assert_close(got[r * b:(r + 1) * b].cpu(), rf.detach()[r * b:(r + 1) * b], tol=2e-5, what="%s window %d" % (name, r))
        if r * b > 0:
            assert torch.isnan(rowstats[1:3, :r * b]).all()          # rows outside the window are not touched
        terms_mean += terms.cpu() / world
        gz = torch.empty(b, D, device=dev())
        gm, gl = torch.empty(B, D, device=dev()), torch.empty(B, D, device=dev())
        N.call("dv_btcvae_bwd_rows", B, D, r * b, b, n_data, 1, N.ptr(rowstats), N.ptr(ws), N.ptr(g_terms), N.ptr(gz), N.ptr(gm),
               N.ptr(gl), N.stream())
        g_z[r * b:(r + 1) * b] = gz.cpu()
        g_mu += gm.cpu()
        g_lv += gl.cpu()
    full = ops.BtcvaeFn.apply(zd, mud, lvd, n_data, True)
    assert_close(terms_mean, full.cpu(), tol=2e-5, what="mean of window terms")
    if not big:
        # gradients of the MEAN-over-windows loss: each window's backward used 1/b, the full batch uses 1/B
        assert_close(g_z / world, zo.grad, what="g_z")
        assert_close(g_mu / world, muo.grad, what="g_mu")
        assert_close(g_lv / world, lvo.grad, what="g_logvar")


def test_u8_to_f32_is_totensor(ops):
    """dv_u8_to_f32 == torchvision ToTensor's `img.float().div(255)` bit for bit (utils/datasets.py:182,247,364-367)."""
    torch.manual_seed(0)
    for n in (16, 1 << 20, 12345 * 16 + 7, 3):
        u = torch.randint(0, 256, (n,), dtype=torch.uint8)
        got = ops.u8_to_f32(u.to(dev())).cpu()
        assert torch.equal(got, u.float().div(255))
    u = torch.arange(256, dtype=torch.uint8).repeat(64)          # every byte value
    assert torch.equal(ops.u8_to_f32(u.to(dev())).cpu(), u.float().div(255))


def test_loss_combine_and_act_bwd_chansum(ops):
    torch.manual_seed(1)
    a = torch.randn(12, requires_grad=True)
    b = torch.randn(3, requires_grad=True)
    ca, cb = [1.0, 0.37], [1.0, 6.0, 0.25]
    ref = ca[0] * a[0] + ca[1] * a[1] + (cb[0] * b[0] + cb[1] * b[1] + cb[2] * b[2])
    (ref * 1.7).backward()
    ad, bd = a.detach().to(dev()).requires_grad_(True), b.detach().to(dev()).requires_grad_(True)
    got = ops.LossCombineFn.apply(ad, bd, ca, cb)
    assert abs(got.item() - ref.item()) <= 1e-6 * abs(ref.item())
    (got * 1.7).backward()
    assert torch.allclose(ad.grad.cpu(), a.grad, rtol=1e-6, atol=0) and torch.allclose(bd.grad.cpu(), b.grad, rtol=1e-6, atol=0)
    # 0-dim second operand (FactorVAE tc), and none at all
    t = torch.tensor(0.8, requires_grad=True)
    td = t.detach().to(dev()).requires_grad_(True)
    got2 = ops.LossCombineFn.apply(ad.detach(), td, [1.0, 1.0], [6.4])
    assert abs(got2.item() - (a[0] + a[1] + 6.4 * t).item()) < 1e-5
    got2.backward()
    assert abs(td.grad.item() - 6.4) < 1e-6 and td.grad.shape == td.shape
    got3 = ops.LossCombineFn.apply(ad.detach(), None, [1.0, 4.0], None)
    assert abs(got3.item() - (a[0] + 4 * a[1]).item()) < 1e-5
    # sigmoid backward + per-channel sums in one pass == the two separate kernels
    for B, C, S in ((37, 3, 64), (5, 1, 32), (300, 1, 64)):
        y = torch.sigmoid(torch.randn(B, C, S, S) * 3).to(dev())
        dy = torch.randn(B, C, S, S).to(dev())
        g, cs = ops.act_bwd_chansum(dy, y, 2)
        g_ref = ops.act_bwd(dy, y, 2)
        assert torch.equal(g, g_ref)
        assert_close(cs.cpu(), g_ref.double().sum((0, 2, 3)).float().cpu(), tol=2e-5, what="chansum")
        _, cs2 = ops.act_bwd_chansum(dy, y, 2)
        assert torch.equal(cs, cs2)


def test_linear_prepacked_equals_per_call_pack(ops):
    """dv_linear_pack_multi (every weight matrix of a node, both operand layouts, one launch) + dv_linear_fwd_packed /
    dv_linear_dgrad_packed give bit-identical results to the per-call packs, incl. the shapes that stay on the CUDA cores."""
    torch.manual_seed(8)
    shapes = [(256, 512), (256, 256), (20, 256), (256, 10), (1000, 1000), (2, 1000), (1000, 10)]
    ws = [(torch.randn(n, k) / math.sqrt(k)).to(dev()) for n, k in shapes]
    packs = ops.linear_pack_multi(ws)
    for (n, k), w, pk in zip(shapes, ws, packs):
        M = 130
        x = torch.randn(M, k, device=dev())
        b = torch.randn(n, device=dev())
        g = torch.randn(M, n, device=dev())
        prev = torch.relu(torch.randn(M, k, device=dev()))
        assert torch.equal(ops.linear_fwd(x, w, b, 1, packed=pk), ops.linear_fwd(x, w, b, 1))
        assert torch.equal(ops.linear_dgrad(g, w, prev, 1, packed=pk), ops.linear_dgrad(g, w, prev, 1))
        assert torch.equal(ops.linear_dgrad(g, w, None, 0, packed=pk), ops.linear_dgrad(g, w, None, 0))


def test_conv_pack_multi_equals_per_layer_pack(ops):
    """dv_conv_pack_multi (all conv layers of a node, one launch) writes exactly the buffers of dv_conv_pack_weights."""
    torch.manual_seed(9)
    chans = [1, 32, 32, 32, 3, 32, 32, 32, 32, 1]        # > 8 layers: the library splits the table
    ws = [(torch.randn(32, ch, 4, 4) * 0.1).to(dev()) for ch in chans]
    for pk, w, ch in zip(ops.conv_pack_multi(ws, chans), ws, chans):
        assert torch.equal(pk, ops.conv_pack(w, ch)), ch
