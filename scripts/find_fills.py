"""Where do the small torch kernels (fill / add / mul) inside one training step come from?  (torch.profiler, eager step)"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disentangling-vae_b200"))
import bench
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
wl = os.environ.get("WORKLOAD", "c2")
trainer = bench.build_job(wl, dev)
_, img, B, *_ = bench.WORKLOADS[wl]
trainer.use_cuda_graph = False
x = torch.rand(B, *img, device=dev)
for i in range(3):
    trainer._step(x, None)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
    trainer._step(x, None)
    torch.cuda.synchronize()
import collections
agg = collections.Counter()
for ev in prof.events():
    if ev.name.startswith("aten::") and ev.name in ("aten::fill_", "aten::zero_", "aten::zeros", "aten::zeros_like", "aten::add", "aten::add_",
                                                     "aten::mul", "aten::cat", "aten::copy_", "aten::sum", "aten::neg", "aten::ones_like"):
        st = [s for s in (ev.stack or []) if "disentangling" in s or "disvae" in s or "autograd" in s][:3]
        agg[(ev.name, str(ev.input_shapes)[:60], " <- ".join(s.split("/")[-1] for s in st))] += 1
for (name, shp, st), n in sorted(agg.items(), key=lambda kv: -kv[1])[:40]:
    print("%3d %-18s %-60s %s" % (n, name, shp, st))
