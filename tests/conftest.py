import json
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, GOLD):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) GPU; run with -m gpu on the GPU box")


@pytest.fixture(scope="session")
def meta():
    with open(os.path.join(GOLD, "meta.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def gold():
    def load(name):
        return {k: v for k, v in np.load(os.path.join(GOLD, name + ".npz")).items()}
    return load


SKIP_KEYS = {"resampler.kernel", "melspectogram_tfmr.spectrogram.window", "melspectogram_tfmr.mel_scale.fb",
             "mfcc_tfmr.dct_mat", "mfcc_tfmr.MelSpectrogram.spectrogram.window", "mfcc_tfmr.MelSpectrogram.mel_scale.fb",
             "sm_anchors", "md_anchors", "lg_anchors", "taper_window"}


@pytest.fixture(scope="session")
def ref_state_dict(meta):
    """Train-form state dict with the deterministic synthetic weights the golden outputs were made with."""
    import synth
    from oracle import ref_port as O
    layout = {k: v for k, v in meta["layout_train"].items() if k not in SKIP_KEYS}
    sd = synth.synth_state_dict(layout, seed=42)
    sd.update(O.frontend_constants())
    for k in ("sm_anchors", "md_anchors", "lg_anchors"):
        sd[k] = torch.tensor(meta["anchors"][k], dtype=torch.float32)
    sd["taper_window"] = torch.empty(0)
    return sd


@pytest.fixture(scope="session")
def cuda_dev():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda", 0)


VARIANTS = {"bottleneck": {"resnet_config": {"block": "Bottleneck"}}, "custom": {"backbone": "custom"},
            "taper": {"taper_input": True}}


@pytest.fixture(scope="session")
def variant_state_dict(meta):
    """(state dict, config) of the other config-selectable backbones with the synthetic weights of
    tests/golden/make_golden_backbones.py; the key layout is the live reference's (backbones_layout.json)."""
    import copy
    import synth
    from oracle import ref_port as O
    with open(os.path.join(GOLD, "backbones_layout.json")) as f:
        layouts = json.load(f)
    cache = {}

    def make(name):
        if name not in cache:
            cfg = copy.deepcopy(O.DEFAULT_CONFIG)
            cfg.update(copy.deepcopy(VARIANTS[name]))
            layout = {k: v for k, v in layouts[name].items() if k not in SKIP_KEYS}
            sd = synth.synth_state_dict(layout, seed=42)
            sd.update(O.frontend_constants())
            for k in ("sm_anchors", "md_anchors", "lg_anchors"):
                sd[k] = torch.tensor(meta["anchors"][k], dtype=torch.float32)
            sd["taper_window"] = torch.empty(0)
            cache[name] = (sd, cfg)
        return cache[name]
    return make
