"""torchrun --nproc-per-node N scripts/ddp_check.py : data-parallel semantics on real GPUs.
Every rank runs ONE Trainer step on its shard of a global batch; checked on rank 0 against the oracle:
  * rank r's loss == oracle loss on shard r
  * post-step parameters == oracle Adam step on the MEAN of the shard gradients
"""
import logging, os, sys, tempfile
from collections import OrderedDict
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "disentangling-vae_b200"))
import disvae
from disvae.models.losses import get_loss_f
from disvae.parallel import broadcast_parameters, shard_batch
from oracle import disvae_oracle as O

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
img, z, per, n_data = (1, 64, 64), 10, 64, 737280
torch.manual_seed(1234 + rank)                     # different init per rank: broadcast must fix it
model = disvae.init_specific_model("Burgess", img, z).to(dev)
broadcast_parameters(model)
p0 = OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items())
opt = torch.optim.Adam(model.parameters(), lr=5e-4)
kw = dict(rec_dist="bernoulli", reg_anneal=0, btcvae_A=1, btcvae_B=6, btcvae_G=1, n_data=n_data)
lf = get_loss_f("btcvae", **kw)
tr = disvae.Trainer(model, opt, lf, device=dev, logger=logging.getLogger("ddp"), save_dir=tempfile.mkdtemp(), is_progress_bar=False)
model.train()
g = torch.Generator().manual_seed(7)
x_global = torch.rand(per * world, *img, generator=g)
eps_global = torch.randn(per * world, z, generator=g)
x = shard_batch(x_global, rank, world); eps = shard_batch(eps_global, rank, world)
model.inject_noise([eps])
loss = tr._train_iteration(x, None)
losses = [None] * world
dist.all_gather_object(losses, loss)
params = OrderedDict((k, v.detach().cpu()) for k, v in model.state_dict().items())
ok = True
if rank == 0:
    leaf = O.make_leaf_params(p0)
    grads = None
    for r in range(world):
        for v in leaf.values(): v.grad = None
        xr, er = shard_batch(x_global, r, world), shard_batch(eps_global, r, world)
        recon, (mu, lv), zz = O.vae_forward(leaf, xr, er)
        l, _ = O.loss_btcvae(xr, recon, mu, lv, zz, n_data, 1, 6, 1, "bernoulli", 1, 0)
        rel = abs(l.item() - losses[r]) / abs(l.item())
        print("rank %d loss %.5f oracle %.5f rel %.2e" % (r, losses[r], l.item(), rel)); ok &= rel < 1e-4
        l.backward()
        gr = [v.grad.clone() for v in leaf.values()]
        grads = gr if grads is None else [a + b for a, b in zip(grads, gr)]
    for v, gsum in zip(leaf.values(), grads): v.grad = gsum / world
    O.make_adam(leaf, 5e-4).step()
    worst = max(((params[k] - leaf[k].detach()).abs().max() / 5e-4).item() for k in params)
    mean = max(((params[k] - leaf[k].detach()).abs().mean() / 5e-4).item() for k in params)
    print("post-step param diff: max %.3f lr, mean %.5f lr" % (worst, mean)); ok &= worst < 2.5 and mean < 0.02
    print("DDP CHECK", "OK" if ok else "FAILED")
# ---- the CUDA-graph data-parallel path (device noise): ranks must stay in lock-step --------------------
torch.manual_seed(99)                               # same Philox key on every rank is fine: shards differ
first = last = None
for it in range(12):
    xb = shard_batch(x_global, rank, world).to(dev)        # a fixed shard: the loss must go down
    v = tr._step(xb, None).item()
    first = v if first is None else first
    last = v
flat = torch.cat([p.detach().flatten() for p in model.parameters()])
gathered = [torch.empty_like(flat) for _ in range(world)]
dist.all_gather(gathered, flat)
in_sync = all(torch.equal(gathered[0], t) for t in gathered)
if rank == 0:
    print("graph path taken:", bool(tr._graphs), "| ranks bit-identical after 12 steps:", in_sync,
          "| loss %.3f -> %.3f" % (first, last))
    ok &= bool(tr._graphs) and in_sync and last < first
    print("DDP GRAPH CHECK", "OK" if ok else "FAILED")
dist.barrier(); dist.destroy_process_group()
sys.exit(0 if ok else 1)
