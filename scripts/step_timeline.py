"""GPU timeline of ONE replayed training step (torch.profiler / CUPTI): every kernel with its stream, start offset, duration
and the idle gap before it on its stream; then the step's span against the per-stream busy time.  WORKLOAD=c2 (default)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disentangling-vae_b200"))
import bench
from torch.profiler import profile, ProfilerActivity
dev = torch.device("cuda", 0)
wl = os.environ.get("WORKLOAD", "c2")
trainer = bench.build_job(wl, dev)
_, img, B, *_ = bench.WORKLOADS[wl]
xs = [torch.rand(B, *img, device=dev) for _ in range(4)]
for i in range(8):                                   # eager warm-up + graph capture + a few replays
    trainer._step(xs[i % 4], None)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for i in range(3):
        trainer._step(xs[i % 4], None)
    torch.cuda.synchronize()
import json, tempfile
path = os.path.join(tempfile.mkdtemp(), "trace.json")
prof.export_chrome_trace(path)
ev = [e for e in json.load(open(path))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
ev.sort(key=lambda e: e["ts"])
n = len(ev) // 3
step = ev[n:2 * n]                                   # the middle one of the three profiled steps
t0 = step[0]["ts"]
last_end, busy = {}, {}
print("%-58s %6s %9s %8s %8s" % ("kernel", "stream", "start us", "dur us", "gap us"))
for e in step:
    st = e.get("args", {}).get("stream", -1)
    s_, d = e["ts"] - t0, e["dur"]
    gap = s_ - last_end[st] if st in last_end else 0.0
    last_end[st] = s_ + d
    busy[st] = busy.get(st, 0.0) + d
    print("%-58s %6s %9.1f %8.1f %8.1f" % (e["name"][:58], st, s_, d, gap))
print("step span %.1f us; busy per stream: %s; kernels %d" % (max(last_end.values()), {k: round(v, 1) for k, v in busy.items()}, len(step)))
