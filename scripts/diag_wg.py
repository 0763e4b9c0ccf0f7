import os, sys, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "disentangling-vae_b200"))
from disvae import ops
dev = torch.device("cuda")
def nhwc(t): return t.permute(0, 2, 3, 1).contiguous()
torch.manual_seed(0)
for B, H in [(1, 16), (4, 8), (2, 16)]:
    x = torch.randn(B, 32, 2 * H, 2 * H); g = torch.randn(B, 32, H, H)
    w = torch.zeros(32, 32, 4, 4, requires_grad=True)
    (F.conv2d(x, w, None, stride=2, padding=1) * g).sum().backward()
    ref = w.grad
    dw, db = ops.conv_wgrad(nhwc(g).to(dev), nhwc(x).to(dev), B, H, H, 32, 0, True)
    dw = dw.cpu(); db = db.cpu()
    print("B,H", B, H, "ref absmax %.3f got absmax %.3f  got nonzero %d/%d" % (ref.abs().max(), dw.abs().max(), (dw != 0).sum(), dw.numel()))
    print("  db err", (db - g.sum((0, 2, 3))).abs().max().item())
    big = ref.abs() > 0.5 * ref.abs().max()
    print("  ratio got/ref on big elems:", (dw[big] / ref[big])[:8].tolist())
    # does got match ref under some permutation of (cl, c, tap)?  correlate
    flat_r, flat_g = ref.flatten(), dw.flatten()
    print("  corr", torch.corrcoef(torch.stack([flat_r, flat_g]))[0, 1].item())
    for name, perm in [("swap cl<->c", ref.permute(1, 0, 2, 3))]:
        print("  corr", name, torch.corrcoef(torch.stack([perm.flatten(), flat_g]))[0, 1].item())
    print("  got[0,0]:", dw[0, 0].flatten()[:8].tolist()); print("  ref[0,0]:", ref[0, 0].flatten()[:8].tolist())
