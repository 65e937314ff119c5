"""Golden fixture for the non-default branches of AudioDetectionLoss, from the LIVE reference (/root/reference, dev container):
cross-entropy class loss (multi_label=False, modules/_loss.py:79-81,157-158) with and without class weights, and the focal
objectness loss (alpha, gamma; modules/_loss.py:9-37,74-77).  Inputs: the predictions / targets of train.npz.

    python tests/golden/make_golden_loss_variants.py      -> tests/golden/loss_variants.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
mp = types.ModuleType("matplotlib"); mp.pyplot = types.ModuleType("matplotlib.pyplot")
sys.modules["matplotlib"] = mp; sys.modules["matplotlib.pyplot"] = mp.pyplot
sys.path.insert(0, REF)
os.chdir(REF)
from modules import AudioDetectionLoss  # noqa: E402
import yaml  # noqa: E402
cfg = yaml.safe_load(open(f"{REF}/config/config.yaml"))
os.chdir(ROOT)

VARIANTS = {
    "ce": dict(multi_label=False),
    "ce_weighted": dict(multi_label=False, class_weights=torch.tensor([0.3, 1.7])),
    "focal": dict(multi_label=True, alpha=0.25, gamma=1.5),
    "focal_ce": dict(multi_label=False, alpha=0.4, gamma=2.0, class_weights=torch.tensor([1.2, 0.6])),
}


def main():
    tr = np.load(os.path.join(HERE, "train.npz"))
    tg = torch.from_numpy(tr["targets"]).clone()
    tg[3, 1] = -100.0          # one ignore_index row
    preds = [torch.from_numpy(tr[f"pred{i}"]) for i in range(3)]
    out = {"targets": tg.numpy()}
    lc = dict(cfg["train_config"]["loss_config"])
    for name, kw in VARIANTS.items():
        c = dict(lc); c.update(kw)
        loss_fn = AudioDetectionLoss(cfg["anchors"], 2, sample_duration=60, **c)
        pg = [p.clone().requires_grad_(True) for p in preds]
        loss, metrics = loss_fn(tuple(pg), tg)
        loss.backward()
        out[f"{name}_loss"] = np.float32(loss.item())
        for i in range(3):
            out[f"{name}_grad{i}"] = pg[i].grad.numpy()
        for k in ("conf_loss", "class_loss", "mean_ciou", "accuracy", "f1"):
            out[f"{name}_{k}"] = np.float64(metrics[k])
        print(name, float(loss), {k: round(float(v), 6) for k, v in metrics.items()})
    np.savez_compressed(os.path.join(HERE, "loss_variants.npz"), **out)


if __name__ == "__main__":
    main()
