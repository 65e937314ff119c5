"""CPU-side contract of the drop-in: state-dict layout, config surface, constants, C-ABI exports, and loud
failure without a GPU (no silent CPU fallback)."""
import copy
import ctypes
import hashlib
import os
import re

import numpy as np
import pytest
import torch
import yaml

import yad_b200
from yad_b200 import _lib, frontend_consts as fc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
torch.set_grad_enabled(False)


def _sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


@pytest.fixture(scope="module")
def model():
    return yad_b200.AudioDetectionNetwork(2).eval()


def test_state_dict_layout_equals_reference(model, meta):
    sd = model.state_dict()
    want = meta["layout_train"]
    assert set(sd) == set(want)
    for k, v in sd.items():
        assert list(v.shape) == want[k], k


def test_deploy_layout_and_fold_equal_reference(meta, ref_state_dict):
    from oracle import ref_port as O
    m = yad_b200.AudioDetectionNetwork(2)
    m.load_state_dict(ref_state_dict)
    m.inference()
    sd = m.state_dict()
    want = meta["layout_deploy"]
    assert set(sd) == set(want)
    folded = O.fold_repvgg(ref_state_dict)
    for k in sd:
        assert list(sd[k].shape) == want[k], k
        if "conv_reparam" in k:
            np.testing.assert_allclose(sd[k].numpy(), folded[k].numpy(), atol=1e-6)
    with pytest.raises(AttributeError):
        m.inference()          # one-shot, like the reference


def test_reference_checkpoint_loads_strict(model, ref_state_dict):
    m = copy.deepcopy(model)
    m.init_zeros_taper_window(ref_state_dict["taper_window"])
    missing = m.load_state_dict(ref_state_dict, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys


def test_frontend_buffers_bit_exact(model, meta):
    sd = model.state_dict()
    for k, h in meta["const_sha256"].items():
        assert _sha(sd[k]) == h, k
    for k in ("sm_anchors", "md_anchors", "lg_anchors"):
        np.testing.assert_array_equal(sd[k].numpy(), np.asarray(meta["anchors"][k], dtype=np.float32))


def test_config_surface(model):
    cfg = yad_b200.load_config()
    for key in ("anchors", "backbone", "block_layers", "resnet_config", "dropout", "melspectrogram_config", "mfcc_config",
                "num_anchors", "train_anchors", "sample_duration", "sample_rate", "new_sample_rate", "scale_input",
                "taper_input", "taper_window", "train_config"):
        assert key in cfg
    assert cfg["train_config"]["optimizer_config"]["weight_decay"] == 0.002
    m2 = yad_b200.AudioDetectionNetwork(3, config=cfg)     # dict config accepted, like the reference
    assert m2.out_channels == 3 * 6
    with pytest.raises(ValueError):
        yad_b200.AudioDetectionNetwork(2, config=7)
    other = copy.deepcopy(cfg); other["backbone"] = "custom"
    assert yad_b200.AudioDetectionNetwork(2, config=other).feature_extractor.fmap4_ch == 1024
    bad = copy.deepcopy(cfg); bad["mfcc_config"]["n_mfcc"] = 20
    with pytest.raises(NotImplementedError):
        yad_b200.AudioDetectionNetwork(2, config=bad)      # unsupported options fail loudly


def test_resample_tap_packing_reconstructs_kernel(model):
    k = model.resampler.kernel
    pk = fc.pack_resample_taps(k)
    P, KW = pk["P"], pk["KW"]
    dense = np.zeros((P, KW + _lib.FE_QW), np.float32)
    taps, base = pk["taps"].numpy(), pk["base"].numpy()
    for u in range(P // 4):
        for i in range(4):
            dense[4 * u + i, base[u]: base[u] + _lib.FE_QW] = taps[u, i]
    np.testing.assert_array_equal(dense[:, :KW], k[:, 0].numpy())
    assert np.all(dense[:, KW:] == 0)
    assert pk["window_len"] == int(base.max()) + _lib.FE_QW


def test_mel_csr_reconstructs_filterbank(model):
    fb = model.melspectogram_tfmr.mel_scale.fb
    csr = fc.pack_mel_csr(fb)
    dense = np.zeros(tuple(fb.shape), np.float32)
    st = csr["start"].numpy()
    for m in range(32):
        dense[csr["bin"].numpy()[st[m]:st[m + 1]], m] = csr["val"].numpy()[st[m]:st[m + 1]]
    np.testing.assert_array_equal(dense, fb.numpy())


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "yad_b200.h")).read()
    declared = set(re.findall(r"\b(yad_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "header parse failed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/yad_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert _lib.load().yad_version() >= 100


def test_no_cpu_fallback(model):
    x = torch.zeros(1, 1, 22050 * 6)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(x, combine_scales=True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        yad_b200.process_model_outputs(torch.zeros(1, 63, 5))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        yad_b200.build_target_by_scale(torch.zeros(3, 4), 120, [1.0, 2.0, 3.0])
    model.train()
    try:
        with pytest.raises((NotImplementedError, RuntimeError)):
            model(x)
    finally:
        model.eval()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "yolo-inspired-audio-activity-detection_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("no oracle", ""), f"{fn} mentions oracle/"


def test_grid_sizes(model):
    from yad_b200.engine import InferenceEngine
    e = InferenceEngine.__new__(InferenceEngine)
    e.rs_P, e.rs_O = 320, 441
    assert InferenceEngine.frames(e, 1323000) == 960
    assert InferenceEngine.grids(e, 1323000) == [120, 60, 30]
    assert InferenceEngine.grids(e, 22050 * 6) == [12, 6, 3]


def test_macro_metrics_match_sklearn():
    """The loss module's accuracy / macro precision / recall / f1 (from the device confusion matrix) vs sklearn, which is what
    the reference calls (modules/_loss.py:167-173)."""
    from sklearn.metrics import accuracy_score, f1_score, precision_score, recall_score
    from yad_b200.train_ops import _macro_metrics
    import warnings
    rng = np.random.default_rng(0)
    for nc, n in ((2, 50), (5, 200), (4, 7)):
        yt, yp = rng.integers(0, nc, n), rng.integers(0, max(1, nc - 1), n)      # some classes never predicted
        cm = np.zeros((nc, nc), np.int64)
        np.add.at(cm, (yt, yp), 1)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            want = (accuracy_score(yt, yp), f1_score(yt, yp, average="macro"), precision_score(yt, yp, average="macro"),
                    recall_score(yt, yp, average="macro"))
        np.testing.assert_allclose(_macro_metrics(cm), want, rtol=1e-12)
