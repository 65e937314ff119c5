"""Shared by the CPU and GPU train-mode tests: the oracle's train-mode forward + loss + autograd on the fixture's inputs."""
import copy

import numpy as np
import torch

import synth
from oracle import ref_port as O

TRAIN_L = 353_000          # as tests/golden/make_golden_train.py: 16 s -> 256 frames -> grids 32 / 16 / 8
TRAIN_CFG = dict(O.DEFAULT_CONFIG, dropout=0.0)


def train_inputs(B=2):
    return synth.synth_clips(B, TRAIN_L, seed=3000), synth.synth_targets(B, seed=21, duration=16.0)


def param_names(sd):
    return [k for k, v in sd.items() if v.dtype.is_floating_point and not k.endswith(("running_mean", "running_var"))
            and (k.endswith((".weight", ".bias")) or k.endswith("_anchors"))]


def oracle_train_step(ref_state_dict, x, tg, dtype=torch.float32):
    """Returns (preds, loss, {name: grad}, updated state dict, x_spectral) from the oracle's train-mode forward and loss.
    The frontend always runs in fp32 (it has no parameters); ``dtype=torch.float64`` runs the CNN, decode, loss and
    autograd in double precision - the yardstick for what fp32 rounding alone does to the deep gradients."""
    xs = O.frontend(x, ref_state_dict)["x_spectral"]
    sd = {k: (v.clone().to(dtype) if v.dtype.is_floating_point else v.clone()) for k, v in ref_state_dict.items()}
    names = param_names(sd)
    L_res = -(-320 * x.shape[-1] // 441)
    with torch.enable_grad():
        for k in names:
            sd[k].requires_grad_(True)
        O._BN_TRAINING[0] = True
        try:
            heads = O.neck(sd, O.backbone(sd, xs.to(dtype), TRAIN_CFG["block_layers"]))
        finally:
            O._BN_TRAINING[0] = False
        preds = O.decode(heads, sd, L_res, xs.shape[-1], 2, TRAIN_CFG, combine_scales=False)
        loss, _ = O.detection_loss(preds, tg.to(dtype), O.DEFAULT_CONFIG["anchors"], 2)
        loss.backward()
    grads = {k: sd[k].grad for k in names}
    return [p.detach() for p in preds], loss.detach(), grads, sd, xs


def grad_stats(g: torch.Tensor) -> np.ndarray:
    g = g.detach().reshape(-1).double().cpu()
    idx = torch.linspace(0, g.numel() - 1, 8).long()
    return np.concatenate([[g.sum().item(), g.norm().item()], g[idx].numpy()])
