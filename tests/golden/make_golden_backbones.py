"""Golden fixtures for the other config-selectable backbones, produced by the LIVE reference
(/root/reference, dev container only):

    python tests/golden/make_golden_backbones.py

  * ``resnet_config.block: Bottleneck`` (modules/_backbone.py:128-138, torchvision Bottleneck, fmaps 256..2048 channels),
  * ``backbone: custom`` (modules/_backbone.py:8-116: 3x7 ExtractorLayers, the height stays 32 so the neck runs in 2-D).

Deterministic synthetic weights (tests/golden/synth.py) are loaded into the reference model, which is run in eval() mode
(train-form and, after ``.inference()``, deploy form) on two seeded 6-second clips.  Stored: heads, predictions and a strided
sample of the first / last feature map -> tests/golden/backbones.npz; the reference's state-dict layouts -> backbones_layout.json."""
import copy
import json
import os
import sys
import types

import numpy as np
import torch
import yaml

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
mp = types.ModuleType("matplotlib"); mp.pyplot = types.ModuleType("matplotlib.pyplot")
sys.modules["matplotlib"] = mp; sys.modules["matplotlib.pyplot"] = mp.pyplot
sys.path.insert(0, REF)
os.chdir(REF)
from modules import AudioDetectionNetwork  # noqa: E402
os.chdir(ROOT)
import synth  # noqa: E402
from make_golden import SKIP  # noqa: E402

torch.set_grad_enabled(False)
VARIANTS = {"bottleneck": {"resnet_config": {"block": "Bottleneck"}}, "custom": {"backbone": "custom"},
            "taper": {"taper_input": True}}       # default ResNet with the clip-long taper window (modules/_architecture.py:87-94)


def variant_config(name):
    cfg = yaml.safe_load(open(f"{REF}/config/config.yaml"))
    cfg.update(copy.deepcopy(VARIANTS[name]))
    return cfg


def main():
    out, layouts = {}, {}
    xs = synth.synth_clips(2, 22050 * 6, seed=3000, silence_tail_every=0)
    for name in VARIANTS:
        m = AudioDetectionNetwork(2, config=variant_config(name))
        layout = {k: list(v.shape) for k, v in m.state_dict().items() if k not in SKIP}
        layouts[name] = {k: list(v.shape) for k, v in m.state_dict().items()}       # (taper_window is still empty here)
        full = dict(m.state_dict()); full.update(synth.synth_state_dict(layout, seed=42))
        m.load_state_dict(full)
        m.eval()
        cap = {}
        m.feature_extractor.register_forward_hook(lambda mod, i, o: cap.update(fmaps=o))
        m.multiscale_module.register_forward_hook(lambda mod, i, o: cap.update(heads=o))
        pt = m(xs, combine_scales=True)
        out[f"{name}.fmap1_s"] = cap["fmaps"][0][:, ::8, ::4, ::2].numpy()
        out[f"{name}.fmap4_s"] = cap["fmaps"][3][:, ::8, ::4, :].numpy()
        for i, h in enumerate(cap["heads"]):
            out[f"{name}.head{i}"] = h.numpy()
        out[f"{name}.preds_train"] = pt.numpy()
        md = copy.deepcopy(m); md.inference()
        out[f"{name}.preds_deploy"] = md(xs, combine_scales=True).numpy()
        print(name, "params", sum(p.numel() for p in m.parameters()), "fmaps", [tuple(f.shape) for f in cap["fmaps"]],
              "preds", tuple(pt.shape), float(pt.sum()))
    np.savez_compressed(os.path.join(HERE, "backbones.npz"), **out)
    with open(os.path.join(HERE, "backbones_layout.json"), "w") as f:
        json.dump(layouts, f)
    print("wrote", os.path.join(HERE, "backbones.npz"), os.path.getsize(os.path.join(HERE, "backbones.npz")))


if __name__ == "__main__":
    main()
