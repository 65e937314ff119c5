"""Golden fixture for the TRAIN-MODE network, made by the LIVE reference
(/root/reference, dev container only):

    python tests/golden/make_golden_train.py

The reference model (train() mode, dropout 0 so that the result is deterministic, synthetic weights of synth.py) runs
two 16-second synthetic clips, the reference's AudioDetectionLoss is back-propagated, and the fixture keeps: the three
prediction tensors, the loss, for EVERY parameter the sum / L2 norm / 8 strided samples of its gradient, the complete
gradient of a few small parameters, and the updated running statistics of three BatchNorm layers."""
import copy
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
mp = types.ModuleType("matplotlib"); mp.pyplot = types.ModuleType("matplotlib.pyplot")
sys.modules["matplotlib"] = mp; sys.modules["matplotlib.pyplot"] = mp.pyplot
sys.path.insert(0, REF)
os.chdir(REF)
import yaml  # noqa: E402
from modules import AudioDetectionNetwork, AudioDetectionLoss  # noqa: E402
os.chdir(ROOT)
import synth  # noqa: E402

SKIP = {"resampler.kernel", "melspectogram_tfmr.spectrogram.window", "melspectogram_tfmr.mel_scale.fb", "mfcc_tfmr.dct_mat",
        "mfcc_tfmr.MelSpectrogram.spectrogram.window", "mfcc_tfmr.MelSpectrogram.mel_scale.fb", "sm_anchors", "md_anchors",
        "lg_anchors", "taper_window"}
TRAIN_L = 353_000          # 16 s at 22.05 kHz -> 256 frames -> grids 32 / 16 / 8
FULL_GRADS = ["feature_extractor.conv1.weight", "feature_extractor.bn1.weight", "feature_extractor.bn1.bias",
              "multiscale_module.rep_block4_1.blocks.0.conv3x3.conv.weight", "multiscale_module.rep_block2_1.conv1.conv1x1.conv.weight",
              "multiscale_module.cspsppf.conv7.conv.bias", "multiscale_module.bic2.conv_c0.norm.weight",
              "sm_anchors", "md_anchors", "lg_anchors"]
STAT_BNS = ["feature_extractor.bn1", "feature_extractor.layer4.1.bn2", "multiscale_module.rep_block3_1.conv1.identity"]


def train_targets(B, seed=21, dur=16.0):
    t = synth.synth_targets(B, seed=seed, duration=dur)
    return t


def main():
    with open(f"{REF}/config/config.yaml") as f:
        cfg = yaml.safe_load(f)
    cfg["dropout"] = 0.0
    torch.manual_seed(42)
    m = AudioDetectionNetwork(2, config=copy.deepcopy(cfg))
    layout = {k: list(v.shape) for k, v in m.state_dict().items() if k not in SKIP}
    full = dict(m.state_dict()); full.update(synth.synth_state_dict(layout, seed=42))
    m.load_state_dict(full)
    m.train()
    x = synth.synth_clips(2, TRAIN_L, seed=3000)
    tg = train_targets(2)
    loss_fn = AudioDetectionLoss(cfg["anchors"], 2, sample_duration=60, **cfg["train_config"]["loss_config"])
    preds = m(x)
    loss, metrics = loss_fn(preds, tg)
    loss.backward()
    out = {"targets": tg.numpy(), "loss": np.float32(loss.item())}
    for i, p in enumerate(preds):
        out[f"pred{i}"] = p.detach().numpy()
    names, stats = [], []
    for k, p in m.named_parameters():
        g = p.grad.reshape(-1).double()
        n = g.numel()
        idx = torch.linspace(0, n - 1, 8).long()
        names.append(k)
        stats.append(np.concatenate([[g.sum().item(), g.norm().item()], g[idx].numpy()]))
    out["grad_names"] = np.array(names)
    out["grad_stats"] = np.stack(stats)
    for k in FULL_GRADS:
        out["grad:" + k] = dict(m.named_parameters())[k].grad.numpy()
    sd = m.state_dict()
    for k in STAT_BNS:
        out["rm:" + k] = sd[k + ".running_mean"].numpy()
        out["rv:" + k] = sd[k + ".running_var"].numpy()
        out["nbt:" + k] = sd[k + ".num_batches_tracked"].numpy()
    out["metrics_keys"] = np.array(sorted(metrics))
    out["metrics_vals"] = np.array([metrics[k] for k in sorted(metrics)], dtype=np.float64)
    np.savez_compressed(os.path.join(HERE, "train_net.npz"), **out)
    print("loss", loss.item(), "size", os.path.getsize(os.path.join(HERE, "train_net.npz")))


if __name__ == "__main__":
    main()
