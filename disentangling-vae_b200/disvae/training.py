"""Training loop with the reference's Trainer API (disvae/training.py:17-196).

Differences from the reference are in HOW, not WHAT: the per-step `loss.item()` host sync
(training.py:164) is replaced by an on-device running sum read once per epoch plus an asynchronous
per-step copy of the loss into pinned host memory (the progress bar shows the latest value that has
already landed), host batches are uploaded one step ahead on a copy stream (`_Prefetcher`), and under
torch.distributed (one process per GPU) gradients are averaged with one flat all-reduce per step
(disvae.parallel).
"""
import logging
import os
from collections import defaultdict
from timeit import default_timer

import torch
from tqdm import trange

from disvae import _native
from disvae.fused import FusedAdam
from disvae.parallel import FlatGradSync, is_distributed
from disvae.utils.modelIO import save_model

TRAIN_LOSSES_LOGFILE = "train_losses.log"


class Trainer():
    """Trainer(model, optimizer, loss_f, device, logger, save_dir, gif_visualizer, is_progress_bar)
    -- training.py:46-62."""

    def __init__(self, model, optimizer, loss_f, device=torch.device("cpu"), logger=logging.getLogger(__name__),
                 save_dir="results", gif_visualizer=None, is_progress_bar=True):
        self.device = device
        self.model = model.to(self.device)
        self.loss_f = loss_f
        self.optimizer = optimizer
        self.save_dir = save_dir
        self.is_progress_bar = is_progress_bar
        self.logger = logger
        self.losses_logger = LossesLogger(os.path.join(self.save_dir, TRAIN_LOSSES_LOGFILE))
        self.gif_visualizer = gif_visualizer
        self.sync_every = 50                      # progress-bar refresh (host sync) period
        self._grad_sync = None
        self._grad_sync_d = None
        self._fused = None                        # FusedAdam over `optimizer` (built lazily on the device)
        self.use_cuda_graph = os.environ.get("DISVAE_CUDA_GRAPH", "1") != "0"
        self._graphs = {}                         # input shape -> (CUDAGraph, static input, static loss)
        self._eligible_steps = 0
        self.logger.info("Training Device: {}".format(self.device))

    def __call__(self, data_loader, epochs=10, checkpoint_every=10):
        """training.py:64-102"""
        start = default_timer()
        self.model.train()
        for epoch in range(epochs):
            storer = defaultdict(list)
            mean_epoch_loss = self._train_epoch(data_loader, storer, epoch)
            self.logger.info('Epoch: {} Average loss per image: {:.2f}'.format(epoch + 1, mean_epoch_loss))
            self.losses_logger.log(epoch, storer)
            if self.gif_visualizer is not None:
                self.gif_visualizer()
            if epoch % checkpoint_every == 0:
                save_model(self.model, self.save_dir, filename="model-{}.pt".format(epoch))
        if self.gif_visualizer is not None:
            self.gif_visualizer.save_reset()
        self.model.eval()
        delta_time = (default_timer() - start) / 60
        self.logger.info('Finished training after {:.1f} min.'.format(delta_time))

    def _train_epoch(self, data_loader, storer, epoch):
        """training.py:104-135; the epoch loss is accumulated on the device."""
        kwargs = dict(desc="Epoch {}".format(epoch + 1), leave=False, disable=not self.is_progress_bar)
        on_gpu = self.device.type == "cuda"
        # own accumulator, updated in place: `loss` may be the CUDA graph's static output tensor, which the next
        # replay overwrites -- never keep a reference to it across steps
        epoch_loss = torch.zeros((), dtype=torch.float32, device=self.device)
        batches = _Prefetcher(data_loader, self.device) if on_gpu else data_loader
        ring = self._loss_ring() if on_gpu else None
        with trange(len(data_loader), **kwargs) as t:
            for i, (data, _) in enumerate(batches):
                loss = self._step(data, storer)
                epoch_loss += loss
                if ring is not None:
                    ring.push(loss)                           # async D2H of this step's loss (4 bytes)
                if self.is_progress_bar and i % self.sync_every == 0:
                    t.set_postfix(loss=ring.latest() if ring is not None else loss.item())
                t.update()
        if self._fused:
            self._fused.flush_state()                         # optimizer.state[p]["step"] follows the device counter
        return epoch_loss.item() / len(data_loader)

    def _loss_ring(self):
        if getattr(self, "_ring", None) is None:
            self._ring = _HostLossRing(self.device)
        return self._ring

    # -- optimizer ---------------------------------------------------------------------------
    def _optimizer_step(self, grad_scale=1.0):
        """Adam through dv_adam_multi when `optimizer` is a plain torch.optim.Adam on CUDA parameters,
        else the optimizer's own step()."""
        if self._fused is None:
            self._fused = FusedAdam(self.optimizer) if FusedAdam.supports(self.optimizer) else False
        if self._fused:
            self._fused.step(grad_scale)
        else:
            self.optimizer.step()

    # -- whole-step CUDA graph ------------------------------------------------------------------
    def _graph_eligible(self, data, storer):
        lf = self.loss_f
        if not (self.use_cuda_graph and self.device.type == "cuda" and self.model.training):
            return False
        if getattr(self.model, "_eps_queue", None):
            return False
        if hasattr(lf, "call_optimize"):
            # FactorVAE: both backward passes and both Adam steps fit one graph in a single process; under data
            # parallelism the graph ends after the second backward pass (gradient average + both Adam steps follow it)
            if getattr(lf, "_perm_queue", None):
                return False
            if is_distributed() and self._grad_sync_d is None:
                return False                                  # the discriminators are broadcast by the first eager step
            if lf._fused_d is None:
                lf._fused_d = FusedAdam(lf.optimizer_d) if FusedAdam.supports(lf.optimizer_d) else False
            if not lf._fused_d:
                return False
        if getattr(lf, "global_batch", False) and is_distributed():
            return False                                      # collectives inside the loss node: run eagerly
        if lf.steps_anneal != 0 and lf.n_train_steps < lf.steps_anneal:
            return False                                      # host-side annealing coefficient still moving
        if storer is not None and (lf.n_train_steps + 1) % lf.record_loss_every == 1:
            return False                                      # this step logs scalars (host sync): run eagerly
        if self._fused is None:
            self._fused = FusedAdam(self.optimizer) if FusedAdam.supports(self.optimizer) else False
        return bool(self._fused)

    def _graph_step(self, data):
        """fwd + loss + bwd (+ Adam when not data-parallel) of one batch as ONE CUDA graph launch (static
        shapes).  Data-parallel: the graph ends after the backward pass; its static gradient tensors are
        gathered into the flat buffer (one kernel), all-reduced (one NCCL call) and consumed by the fused
        Adam launch with grad_scale = 1/world."""
        key = (tuple(data.shape), str(data.dtype))
        entry = self._graphs.get(key)
        ddp = is_distributed()
        if entry is None:
            static_x = torch.empty(data.shape, dtype=torch.float32, device=self.device)
            self._fill_static(static_x, data)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            steps_before = self.loss_f.n_train_steps
            launches_before = _native.lib().dv_launch_count()
            factor = hasattr(self.loss_f, "call_optimize")
            with torch.cuda.graph(g):
                if factor:
                    with torch.no_grad():
                        self.model(static_x)                      # the discarded full-batch forward of training.py:153:
                                                                  # kept for its noise draw (same stream as the eager path)
                    if ddp:                                       # both backward passes; the optimizers step after the average
                        loss = self.loss_f.call_optimize(static_x, self.model, self.optimizer, None, step_optimizers=False)
                    else:
                        loss = self.loss_f.call_optimize(static_x, self.model, _StepProxy(self.optimizer, self._optimizer_step), None)
                        self._fused.host_steps -= 1               # capture executed nothing
                        self.loss_f._fused_d.host_steps -= 1
                else:
                    recon, latent_dist, z = self.model(static_x)
                    loss = self.loss_f(static_x, recon, latent_dist, True, None, latent_sample=z)
                    self.optimizer.zero_grad(set_to_none=True)
                    loss.backward()
                    if not ddp:
                        self._fused.step()
                        self._fused.host_steps -= 1               # capture executed nothing
                static_loss = loss.detach()
            self.loss_f.n_train_steps = steps_before
            flat = None
            if ddp:
                params = [p for p in self.model.parameters() if p.grad is not None]
                if factor:                                        # one flat buffer (one all-reduce) for both networks
                    params += [p for p in self.loss_f.discriminator.parameters() if p.grad is not None]
                static_grads = [p.grad for p in params]           # written by every replay
                flat_buf = torch.zeros(sum(t.numel() for t in static_grads), dtype=torch.float32, device=self.device)
                off, views = 0, []
                for p in params:                                  # Adam reads the all-reduced flat views
                    views.append(flat_buf[off:off + p.numel()].view_as(p))
                    off += p.numel()
                flat = (flat_buf, [t.view(-1) for t in static_grads], params, views)
            entry = (g, static_x, static_loss, _native.lib().dv_launch_count() - launches_before, flat)
            self._graphs[key] = entry
        g, static_x, static_loss, n_kernels, flat = entry
        self._fill_static(static_x, data)
        g.replay()
        _native.GRAPH_LAUNCHES += n_kernels
        if flat is not None:
            import torch.distributed as dist
            torch.cat(flat[1], out=flat[0])
            dist.all_reduce(flat[0], op=dist.ReduceOp.SUM)
            for p, v in zip(flat[2], flat[3]):                    # (an eager step in between re-binds .grad)
                p.grad = v
            self._fused.step(grad_scale=1.0 / dist.get_world_size())
            self._fused.host_steps -= 1
            if hasattr(self.loss_f, "call_optimize"):
                self.loss_f._fused_d.step(grad_scale=1.0 / dist.get_world_size())
                self.loss_f._fused_d.host_steps -= 1
        self.loss_f.n_train_steps += 1
        self._fused.host_steps += 1
        if getattr(self.loss_f, "_fused_d", None):
            self.loss_f._fused_d.host_steps += 1
        return static_loss

    def _fill_static(self, static_x, data):
        """The graph's input buffer <- this step's batch.  uint8 batches are uploaded as bytes and
        converted by dv_u8_to_f32 straight into the buffer (ToTensor's /255), float batches are copied."""
        if data.dtype == torch.uint8:
            from disvae import ops
            ops.u8_to_f32(data.to(self.device, non_blocking=True), out=static_x)
        else:
            static_x.copy_(data, non_blocking=True)

    def _step(self, data, storer):
        """One optimisation step; returns the loss as a detached 0-dim device tensor."""
        if self._graph_eligible(data, storer):
            self._eligible_steps += 1
            if self._eligible_steps > 2 or (tuple(data.shape), str(data.dtype)) in self._graphs:   # 2 eager warm-up steps first
                return self._graph_step(data)
        data = data.to(self.device, non_blocking=True)
        if data.dtype == torch.uint8:                         # bytes over PCIe, ToTensor's /255 on the device
            from disvae import ops
            data = ops.u8_to_f32(data)
        recon_batch, latent_dist, latent_sample = self.model(data)
        try:
            loss = self.loss_f(data, recon_batch, latent_dist, self.model.training, storer, latent_sample=latent_sample)
        except ValueError:
            # losses with several optimizers (FactorVAE) announce themselves by raising from __call__
            # (training.py:160-162, losses.py:240-241).  Only the loss call is inside the `try`: a ValueError from
            # anywhere else in the step (a kernel-side shape check, the optimizer) must surface, not be rerouted.
            if not hasattr(self.loss_f, "call_optimize"):
                raise
            if is_distributed():
                loss = self._factor_step_distributed(data, storer)
            else:
                loss = self.loss_f.call_optimize(data, self.model, _StepProxy(self.optimizer, self._optimizer_step), storer)
            return loss.detach()
        self.optimizer.zero_grad()
        loss.backward()
        self._sync_grads()
        self._optimizer_step()
        return loss.detach()

    def _train_iteration(self, data, storer):
        """training.py:137-164 (returns a Python float, i.e. synchronises)."""
        return self._step(data, storer).item()

    # -- data parallel ---------------------------------------------------------------------
    def _sync_grads(self):
        if not is_distributed():
            return
        if self._grad_sync is None:
            self._grad_sync = FlatGradSync(list(self.model.parameters()))
        self._grad_sync.sync()

    def _factor_step_distributed(self, data, storer):
        """FactorVAE step with the optimizer updates deferred until gradients are averaged."""
        loss = self._factor_grads_distributed(data, storer)
        self._optimizer_step()
        self.loss_f._step_d()
        return loss

    def _factor_grads_distributed(self, data, storer, **inject):
        """Both backward passes of losses.py:243-313 on this rank's shard, then the rank-average of the VAE and the
        discriminator gradients (two flat all-reduces); no optimizer steps."""
        lf = self.loss_f
        if self._grad_sync_d is None:
            from disvae.parallel import broadcast_parameters
            broadcast_parameters(lf.discriminator)                # replicas must start from identical discriminators
            self._grad_sync_d = FlatGradSync(list(lf.discriminator.parameters()))
        loss = lf.call_optimize(data, self.model, self.optimizer, storer, step_optimizers=False, **inject)
        self._sync_grads()
        self._grad_sync_d.sync()
        return loss

    def _grads_only(self, data, storer=None, **inject):
        """Forward + loss + backward (+ the data-parallel gradient average) of one batch WITHOUT an optimizer step:
        afterwards every `p.grad` (and, for FactorVAE, the discriminator's) holds exactly what the optimizers would
        consume.  Used by the parity checks (bench.py `parity` / `ddp_parity`, tests); `inject` forwards
        eps1/eps2/perms to FactorKLoss.call_optimize."""
        data = data.to(self.device, non_blocking=True)
        lf = self.loss_f
        if hasattr(lf, "call_optimize"):
            if is_distributed():
                return self._factor_grads_distributed(data, storer, **inject).detach()
            return lf.call_optimize(data, self.model, self.optimizer, storer, step_optimizers=False, **inject).detach()
        recon_batch, latent_dist, latent_sample = self.model(data)
        loss = lf(data, recon_batch, latent_dist, self.model.training, storer, latent_sample=latent_sample)
        self.optimizer.zero_grad()
        loss.backward()
        self._sync_grads()
        return loss.detach()


class _Prefetcher:
    """Iterates a loader of (data, label) batches one step ahead: the host->device copy of batch i+1 is issued
    on a copy stream while step i runs (two device buffers, reused).  Pageable host tensors still work (the copy
    is then synchronous); batches already on the device pass through."""

    def __init__(self, loader, device):
        self.loader, self.device = loader, device
        self.stream = torch.cuda.Stream(device)
        self.bufs = [None, None]
        self.consumed = [None, None]              # main-stream events: buffer may be overwritten after this

    def _upload(self, item, slot):
        data, label = item
        if not torch.is_tensor(data) or data.device.type == "cuda":
            return data, label, None
        buf = self.bufs[slot]
        if buf is None or buf.shape != data.shape or buf.dtype != data.dtype:
            buf = self.bufs[slot] = torch.empty(data.shape, dtype=data.dtype, device=self.device)
        with torch.cuda.stream(self.stream):
            if self.consumed[slot] is not None:
                self.stream.wait_event(self.consumed[slot])
            buf.copy_(data, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        return buf, label, ready

    def __iter__(self):
        it = iter(self.loader)
        slot = 0
        try:
            nxt = self._upload(next(it), slot)
        except StopIteration:
            return
        while nxt is not None:
            data, label, ready = nxt
            cur_slot = slot
            slot ^= 1
            try:
                nxt = self._upload(next(it), slot)            # overlaps with the step on `data`
            except StopIteration:
                nxt = None
            main = torch.cuda.current_stream(self.device)
            if ready is not None:
                main.wait_event(ready)
            yield data, label
            if ready is not None:                             # the step that read `data` has been enqueued
                ev = torch.cuda.Event()
                ev.record(main)
                self.consumed[cur_slot] = ev

    def __len__(self):
        return len(self.loader)


class _HostLossRing:
    """Per-step losses copied asynchronously into pinned host memory; `latest()` returns the newest value that
    has already arrived (never blocks)."""

    def __init__(self, device, size=64):
        self.host = torch.zeros(size, dtype=torch.float32).pin_memory()
        self.events = [None] * size
        self.n = 0
        self.size = size

    def push(self, loss):
        k = self.n % self.size
        self.host[k:k + 1].copy_(loss.detach().reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        self.events[k] = ev
        self.n += 1

    def latest(self):
        for back in range(1, min(self.n, self.size) + 1):
            k = (self.n - back) % self.size
            if self.events[k] is not None and self.events[k].query():
                return float(self.host[k])
        return float("nan")


class _StepProxy:
    """What FactorKLoss.call_optimize sees as `optimizer`: zero_grad() of the real one, step() through
    the Trainer's fused Adam."""

    def __init__(self, optimizer, step_fn):
        self._opt, self._step = optimizer, step_fn

    def zero_grad(self, *a, **k):
        return self._opt.zero_grad(*a, **k)

    def step(self):
        return self._step()


class LossesLogger(object):
    """CSV "Epoch,Loss,Value" writer (training.py:167-190)."""

    def __init__(self, file_path_name):
        if os.path.isfile(file_path_name):
            os.remove(file_path_name)
        self.logger = logging.getLogger("losses_logger")
        self.logger.setLevel(1)
        file_handler = logging.FileHandler(file_path_name)
        file_handler.setLevel(1)
        self.logger.addHandler(file_handler)
        self.logger.debug(",".join(["Epoch", "Loss", "Value"]))

    def log(self, epoch, losses_storer):
        for k, v in losses_storer.items():
            self.logger.debug(",".join(str(item) for item in [epoch, k, mean(v)]))


def mean(l):
    return sum(l) / len(l)
