import os, sys, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "disentangling-vae_b200"))
from disvae import ops
dev = torch.device("cuda")
def nhwc(t): return t.permute(0, 2, 3, 1).contiguous()
def nchw(t): return t.permute(0, 3, 1, 2).contiguous()
torch.manual_seed(0)
H = int(os.environ.get("HH", "16"))
for B in [100, 148, 170, 197, 200, 250, 300, 400]:
    lo = torch.randn(B, 32, H, H); w = torch.randn(32, 32, 4, 4) * 0.1
    ref = F.conv_transpose2d(lo, w, None, stride=2, padding=1)
    wp = ops.conv_pack(w.to(dev), 32); x = nhwc(lo).to(dev)
    errs = []
    for rep in range(6):
        got = nchw(ops.conv_up(x, wp, None, None, B, H, H, 32, 0, 0).cpu())
        d = (got - ref).abs()
        e = (d.max() / ref.abs().max()).item()
        bad = (d.amax(dim=(1, 2, 3)) > 1e-3 * ref.abs().max()).nonzero().flatten().tolist()
        errs.append((round(e, 6), bad[:6], len(bad)))
    print("B=%d tiles=%d" % (B, B * ((H + 6) // 7) if H == 16 else -1), errs)
