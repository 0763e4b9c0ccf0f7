"""Flip-robust full-batch gradient parity.  TEST INFRASTRUCTURE (tests/test_fullsize_gpu.py, tests/test_model_gpu.py,
bench.py's `parity` / `ddp_parity` checker).

Problem: at the batch sizes BASELINE.json quotes (1024 / 512 / 256 images = 10^7..10^8 ReLU units) two CORRECT fp32
evaluations of the Burgess VAE round a handful of pre-activations to opposite sides of zero; each such unit switches a
back-propagated path on or off, so weight gradients of two correct implementations differ by 1e-3..1e-1 of their scale
(the CPU fp32 oracle against its own fp64 run: 9e-4 .. 7e-2 on the four bench workloads) -- far above any arithmetic
error and useless as a referee.

Method: (1) run the CUDA path with ops.start_trace(): it reports the on/off pattern of every (Leaky)ReLU it evaluated;
(2) run the oracle in fp64 recording its pre-activations: every unit where the CUDA pattern disagrees with the fp64 sign
must be numerically AMBIGUOUS (|pre| tiny against the layer's scale) -- otherwise the forward pass is wrong;
(3) run the fp64 oracle again with the CUDA pattern imposed (pre * mask instead of its own sign test): its backward
pass is now a smooth function evaluated on the same branch of the network, and the CUDA gradients must match it to
1e-4 of every tensor's scale.  (2)+(3) together are the full-batch gradient statement: same function, same branch,
fp32-grade arithmetic."""
from collections import OrderedDict

import torch

from oracle import disvae_oracle as O  # noqa: E402

ENC_CONVS = ("conv1", "conv2", "conv3", "conv_64")
DEC_CONVTS = ("convT_64", "convT1", "convT2")


def _to_oracle_names(trace, params):
    """[(product trace name, tensor)] in call order -> {oracle activation name: bool mask in the oracle's layout}."""
    enc_names = [n for n in ENC_CONVS if ("encoder.%s.weight" % n) in params]
    dec_names = [n for n in DEC_CONVTS if ("decoder.%s.weight" % n) in params]
    calls = {"encoder": -1, "decoder": -1, "disc": -1}
    masks = OrderedDict()
    for name, t in trace:
        mod, layer = name.split(".")
        t = t.detach()
        if mod == "encoder":
            if layer == "conv0":
                calls["encoder"] += 1
            tag = "encoder#%d." % calls["encoder"]
            if layer.startswith("conv"):
                masks[tag + enc_names[int(layer[4:])]] = (t > 0).permute(0, 3, 1, 2).cpu()      # NHWC -> NCHW
            else:
                masks[tag + layer] = (t > 0).cpu()
        elif mod == "decoder":
            if layer == "lin1":
                calls["decoder"] += 1
            tag = "decoder#%d." % calls["decoder"]
            if layer.startswith("convT"):
                masks[tag + dec_names[int(layer[5:])]] = (t > 0).permute(0, 3, 1, 2).cpu()
            else:
                masks[tag + layer] = (t > 0).cpu()
        elif mod == "mlp":
            if layer == "lin1":
                calls["disc"] += 1
            masks["disc#%d.%s" % (calls["disc"], layer)] = (t > 0).cpu()
    return masks


def same_branch_reference(trace, params32, run_oracle, disc32=None, run_oracle32=None):
    """`run_oracle(p, dp)` evaluates the oracle's loss in fp64 on leaf params `p` (and discriminator `dp`, or None),
    runs its backward pass(es) and returns the loss.
    -> dict(loss, grads {name: fp64 grad}, flips, units, flip_max_rel = largest |pre-activation| among the units where
    the CUDA on/off pattern differs from the fp64 sign, relative to the layer's mean |pre-activation|)."""
    masks = _to_oracle_names(trace, params32)
    p64 = OrderedDict((k, v.detach().cpu().double()) for k, v in params32.items())
    d64 = OrderedDict((k, v.detach().cpu().double()) for k, v in disc32.items()) if disc32 is not None else None
    # ONE fp64 pass that records every pre-activation AND continues on the CUDA path's branch (pre * mask)
    O._ACT = dict(record={}, masks=masks, calls={})
    try:
        p = O.make_leaf_params(p64)
        dp = O.make_leaf_params(d64) if d64 is not None else None
        loss = run_oracle(p, dp)
        rec = O._ACT["record"]
    finally:
        O._ACT = None
    assert set(rec) == set(masks), (sorted(rec), sorted(masks))
    flips, units, worst = 0, 0, 0.0
    for k, pre in rec.items():
        m = masks[k]
        assert tuple(m.shape) == tuple(pre.shape), (k, m.shape, pre.shape)
        dis = m != (pre > 0)
        units += m.numel()
        n = int(dis.sum())
        if n:
            flips += n
            scale = pre.abs().mean().clamp_min(1e-30)
            worst = max(worst, (pre[dis].abs().max() / scale).item())
    grads = OrderedDict((k, v.grad) for k, v in p.items())
    if dp is not None:
        grads.update(("disc." + k, v.grad) for k, v in dp.items())
    out = dict(loss=float(loss), grads=grads, flips=flips, units=units, flip_max_rel=worst)
    if run_oracle32 is not None:
        # calibration: the CPU fp32 oracle on the SAME branch -- what plain fp32 arithmetic (MKL/oneDNN) loses against
        # fp64 when no unit flips; the CUDA path (error-compensated 3xTF32, ~5x the rounding of an fp32 FMA chain per
        # kernel) is judged relative to it
        O._ACT = dict(record=None, masks=masks, calls={})
        try:
            q = O.make_leaf_params(OrderedDict((k, v.detach().cpu().float()) for k, v in params32.items()))
            dq = O.make_leaf_params(OrderedDict((k, v.detach().cpu().float()) for k, v in disc32.items())) if disc32 is not None else None
            run_oracle32(q, dq)
        finally:
            O._ACT = None
        g32 = OrderedDict((k, v.grad) for k, v in q.items())
        if dq is not None:
            g32.update(("disc." + k, v.grad) for k, v in dq.items())
        out["cpu_fp32_same_branch_err"], out["cpu_fp32_worst_tensor"] = grad_errors(g32, grads)
    return out


def grad_errors(ours, ref):
    """max over tensors of max|a - b| / scale, and the tensor where it happens.  scale = the largest |entry| among the
    fp64 gradients of the tensor's LAYER (weight and bias together): a bias gradient is a sum of the same upstream
    gradients that, times O(1) activations, make up the weight gradient -- but it can cancel to (almost) nothing (the two
    logits' biases of the FactorVAE discriminator at initialisation: 512 terms of +-1e-3 summing to 1e-5), and judging its
    rounding error against its own magnitude measures conditioning, not arithmetic."""
    layer_scale = {}
    for k, b in ref.items():
        layer = k.rsplit(".", 1)[0]
        layer_scale[layer] = max(layer_scale.get(layer, 0.0), b.detach().abs().max().item())
    worst, key = 0.0, None
    for k, b in ref.items():
        a = ours[k].detach().double().cpu()
        e = ((a - b.detach().double().cpu()).abs().max().item()) / max(layer_scale[k.rsplit(".", 1)[0]], 1e-30)
        if e > worst:
            worst, key = e, k
    return worst, key
