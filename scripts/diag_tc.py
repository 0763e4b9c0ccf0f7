import os, sys, torch, torch.nn.functional as F
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "disentangling-vae_b200"))
from disvae import ops
dev = torch.device("cuda")
def nhwc(t): return t.permute(0, 2, 3, 1).contiguous()
def nchw(t): return t.permute(0, 3, 1, 2).contiguous()
def tf32_exact(t): return (t.view(torch.int32) & -8192).view(torch.float32)
torch.manual_seed(0)
B, H = 8, 16
for name, fa, fw in [("a,w arbitrary", lambda t: t, lambda t: t), ("a tf32-exact", tf32_exact, lambda t: t),
                     ("w tf32-exact", lambda t: t, tf32_exact), ("both exact", tf32_exact, tf32_exact)]:
    x = fa(torch.randn(B, 32, 2 * H, 2 * H)); w = fw(torch.randn(32, 32, 4, 4) * 0.1); b = torch.randn(32)
    ref = F.conv2d(x.double(), w.double(), b.double(), stride=2, padding=1)
    wp = ops.conv_pack(w.to(dev), 32)
    got = nchw(ops.conv_down(nhwc(x).to(dev), wp, b.to(dev), None, B, H, H, 32, 0, 0).cpu()).double()
    f32 = F.conv2d(x, w, b, stride=2, padding=1).double()
    print("%-16s tc err/max %.3e   fp32-cpu err/max %.3e" % (name, ((got - ref).abs().max() / ref.abs().max()).item(),
                                                            ((f32 - ref).abs().max() / ref.abs().max()).item()))
# timing at the bench shape (conv2: B=1024, lo 16x16)
B, H = 1024, 16
x = torch.randn(B, 2 * H, 2 * H, 32, device=dev); w = torch.randn(32, 32, 4, 4, device=dev) * 0.1
wp = ops.conv_pack(w, 32)
for _ in range(3): ops.conv_down(x, wp, None, None, B, H, H, 32, 0, 1)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): ops.conv_down(x, wp, None, None, B, H, H, 32, 0, 1)
e1.record(); torch.cuda.synchronize()
print("conv2-shape down: %.1f us" % (e0.elapsed_time(e1) * 100))
