"""A few training steps of the bench workload (for ncu captures)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disentangling-vae_b200"))
import bench
dev = torch.device("cuda", 0)
wl = os.environ.get("WORKLOAD", "c2")
trainer = bench.build_job(wl, dev)
img, B = bench.WORKLOADS[wl][1], bench.WORKLOADS[wl][2]
x = torch.rand(B, *img, device=dev)
for i in range(int(os.environ.get("STEPS", "3"))):
    trainer._step(x, None)
torch.cuda.synchronize()
print("done")
