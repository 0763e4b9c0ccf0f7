"""A few training steps of the bench workload (for ncu captures)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disentangling-vae_b200"))
import bench
dev = torch.device("cuda", 0)
trainer, (loss_name, img, B, z, n_data, lr) = bench.build_job(os.environ.get("WORKLOAD", "c2"), dev)
x = torch.rand(B, *img, device=dev)
for i in range(int(os.environ.get("STEPS", "3"))):
    trainer._step(x, None)
torch.cuda.synchronize()
print("done")
