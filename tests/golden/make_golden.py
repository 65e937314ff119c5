"""Generates the golden fixtures by running the LIVE reference (/root/reference, dev container only).

    python tests/golden/make_golden.py

Imports the reference with a matplotlib stub, loads the deterministic synthetic weights of
tests/golden/synth.py into it, runs it on seeded synthetic clips and stores the outputs as small .npz
fixtures next to this script.  The GPU box has no /root/reference: tests only read the fixtures."""
import copy
import hashlib
import json
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
mp = types.ModuleType("matplotlib"); mp.pyplot = types.ModuleType("matplotlib.pyplot")
sys.modules["matplotlib"] = mp; sys.modules["matplotlib.pyplot"] = mp.pyplot
sys.path.insert(0, REF)
os.chdir(REF)
from modules import AudioDetectionNetwork, AudioDetectionLoss  # noqa: E402
from dataset import AudioDataset  # noqa: E402
from smoothener import EMAParamsSmoothener  # noqa: E402
import inference as ref_inf  # noqa: E402  (re-seeds torch to 42 on import)
import torchvision  # noqa: E402
os.chdir(ROOT)
import synth  # noqa: E402

torch.set_grad_enabled(False)
CONST_KEYS = ["resampler.kernel", "melspectogram_tfmr.spectrogram.window", "melspectogram_tfmr.mel_scale.fb",
              "mfcc_tfmr.dct_mat", "mfcc_tfmr.MelSpectrogram.spectrogram.window", "mfcc_tfmr.MelSpectrogram.mel_scale.fb"]
SKIP = set(CONST_KEYS) | {"sm_anchors", "md_anchors", "lg_anchors", "taper_window"}


def sha(t):
    return hashlib.sha256(t.contiguous().numpy().tobytes()).hexdigest()


def build_model():
    m = AudioDetectionNetwork(2, config=f"{REF}/config/config.yaml")
    layout = {k: list(v.shape) for k, v in m.state_dict().items() if k not in SKIP}
    sd = synth.synth_state_dict(layout, seed=42)
    full = dict(m.state_dict()); full.update(sd)
    m.load_state_dict(full)
    return m.eval(), layout


def ref_keep(out, iou):
    """inference.py:55-84 verbatim up to `keep` (the reference does not return it)."""
    cw = out[..., -2:]
    x1 = cw[..., :1] - (cw[..., -1:] / 2); x2 = cw[..., :1] + (cw[..., -1:] / 2)
    y1 = torch.zeros_like(x1); y2 = torch.zeros_like(x2) + 10
    coords = torch.cat([x1, y1, x2, y2], dim=-1).clip(min=0, max=60).squeeze(0)
    obj = out[..., :1].sigmoid()
    cs = torch.nn.functional.softmax(out[..., 1:-2], dim=-1)
    cs = torch.gather(cs, dim=-1, index=cs.argmax(dim=-1, keepdim=True))
    conf = cs * obj
    B, P = out.shape[0], out.shape[1]
    bidx = torch.arange(B).repeat_interleave(P)
    keep = torchvision.ops.batched_nms(coords.flatten(0, -2), conf.reshape(-1), idxs=bidx, iou_threshold=iou)
    return keep, conf.reshape(B, P), coords.reshape(B, P, 4)


def main():
    m, layout = build_model()
    sd = m.state_dict()
    md = copy.deepcopy(m); md.inference()
    meta = {
        "layout_train": {k: list(v.shape) for k, v in sd.items()},
        "layout_deploy": {k: list(v.shape) for k, v in md.state_dict().items()},
        "const_sha256": {k: sha(sd[k]) for k in CONST_KEYS},
        "const_sum": {k: float(sd[k].double().sum()) for k in CONST_KEYS},
        "anchors": {k: sd[k].tolist() for k in ("sm_anchors", "md_anchors", "lg_anchors")},
        "versions": {"torch": torch.__version__, "torchvision": torchvision.__version__},
    }
    cap = {}
    m.feature_extractor.register_forward_hook(lambda mod, i, o: cap.update(x_spectral=i[0], fmaps=o))
    m.multiscale_module.register_forward_hook(lambda mod, i, o: cap.update(heads=o))
    capd = {}
    md.multiscale_module.register_forward_hook(lambda mod, i, o: capd.update(heads=o))

    # ---- (a) short clips: 6 s -> 96 frames, B = 3 (clip 2 has a silence tail... every 3rd here)
    xs = synth.synth_clips(3, 22050 * 6, seed=1000, silence_tail_every=3)
    pre = {}
    r16 = m.resampler(xs)
    mel = m.melspectogram_tfmr(r16)
    mfcc = m.mfcc_tfmr(r16)
    out_t = m(xs, combine_scales=True)
    out_d = md(xs, combine_scales=True)
    short = {
        "resampled_head": r16[:, 0, :4000].numpy(), "mel": mel.numpy(), "mfcc": mfcc.numpy(),
        "x_spectral": cap["x_spectral"].numpy(),
        "fmap1": cap["fmaps"][0].numpy(), "fmap4": cap["fmaps"][3].numpy(),
        "head_sm": cap["heads"][0].numpy(), "head_md": cap["heads"][1].numpy(), "head_lg": cap["heads"][2].numpy(),
        "preds_train": out_t.numpy(), "preds_deploy": out_d.numpy(),
        "dhead_sm": capd["heads"][0].numpy(), "dhead_md": capd["heads"][1].numpy(), "dhead_lg": capd["heads"][2].numpy(),
    }
    np.savez_compressed(os.path.join(HERE, "short_clips.npz"), **short)

    # ---- (b) one full-length 60 s clip (config 1 of BASELINE.json)
    xf = synth.synth_clips(1, 1323000, seed=2000, silence_tail_every=0)
    out_f = m(xf, combine_scales=True)
    out_fd = md(xf, combine_scales=True)
    seg, bidx = ref_inf.process_model_outputs(out_f.clone(), 0.1, 0.2)
    full = {"x_spectral": cap["x_spectral"].numpy(), "preds_train": out_f.numpy(), "preds_deploy": out_fd.numpy(),
            "seg_0.1_0.2": seg.numpy(), "bidx_0.1_0.2": bidx.numpy()}
    np.savez_compressed(os.path.join(HERE, "full_clip.npz"), **full)

    # ---- (c) NMS / post-processing on planted head tensors
    nms = {}
    for name, B in (("b1", 1), ("b3", 3)):
        o = synth.synth_heads(B, 630, 2, seed=7 + B)
        for iou, cthr in ((0.1, 0.2), (0.1, 0.65), (0.05, 0.5)):
            keep, conf, coords = ref_keep(o, iou)
            tag = f"{name}_{iou}_{cthr}"
            nms[f"keep_{tag}"] = keep.numpy()
            seg, bidx = ref_inf.process_model_outputs(o.clone(), iou, cthr)
            nms[f"seg_{tag}"] = seg.numpy(); nms[f"bidx_{tag}"] = bidx.numpy()
        nms[f"conf_{name}"] = conf.numpy(); nms[f"coords_{name}"] = coords.numpy()
    # also the model's own output (realistic score distribution)
    keep, conf, coords = ref_keep(out_t, 0.1)
    nms["keep_model_0.1"] = keep.numpy()
    for iou, cthr in ((0.1, 0.2), (0.05, 0.1)):
        seg, bidx = ref_inf.process_model_outputs(out_t.clone(), iou, cthr)
        nms[f"seg_model_{iou}_{cthr}"] = seg.numpy(); nms[f"bidx_model_{iou}_{cthr}"] = bidx.numpy()
    try:
        ref_inf.process_model_outputs(out_t.clone(), 0.1, 0.999999)
        meta["empty_raises"] = "no"
    except Exception as e:  # noqa: BLE001
        meta["empty_raises"] = type(e).__name__
    np.savez_compressed(os.path.join(HERE, "nms.npz"), **nms)

    # ---- (d) anchor matching, (e) loss, (f) EMA / Adam
    cfg = m.config
    tg = synth.synth_targets(4, seed=11)
    tr = {"targets": tg.numpy()}
    for name, G in (("sm", 120), ("md", 60), ("lg", 30)):
        (bi, gi, ai), cl, cw = AudioDataset.build_target_by_scale(tg, G, cfg["anchors"][name], anchor_threshold=5,
                                                                  sample_duration=60, edge_threshold=0.5)
        tr[f"bi_{name}"], tr[f"gi_{name}"], tr[f"ai_{name}"] = bi.numpy(), gi.numpy(), ai.numpy()
        tr[f"cl_{name}"], tr[f"cw_{name}"] = cl.numpy(), cw.numpy()
    g = torch.Generator().manual_seed(5)
    preds = []
    for G in (120, 60, 30):
        p = torch.randn(4, G, 3, 5, generator=g)
        p[..., 3] = torch.rand(4, G, 3, generator=g) * 60; p[..., 4] = torch.rand(4, G, 3, generator=g) * 40
        preds.append(p)
    lc = cfg["train_config"]["loss_config"]
    loss_fn = AudioDetectionLoss(cfg["anchors"], 2, sample_duration=60, **lc)
    with torch.enable_grad():
        pg = [p.clone().requires_grad_(True) for p in preds]
        loss, metrics = loss_fn(tuple(pg), tg)
        loss.backward()
    tr["loss"] = np.float32(loss.item())
    for i, p in enumerate(preds):
        tr[f"pred{i}"] = p.numpy(); tr[f"grad{i}"] = pg[i].grad.numpy()
    meta["loss_metrics"] = {k: float(v) for k, v in metrics.items()}
    # Adam (L2 wd) + EMA on a small arena: 3 steps
    gp = torch.Generator().manual_seed(9)
    w = torch.randn(1000, generator=gp)
    lin = nn.Linear(1, 1); par = nn.Parameter(w.clone())
    oc = cfg["train_config"]["optimizer_config"]
    opt = torch.optim.Adam([par], lr=oc["lr"], betas=tuple(oc["betas"]), eps=oc["eps"], weight_decay=oc["weight_decay"])
    ema = w.clone()
    grads = []
    for step in range(1, 4):
        gr = torch.randn(1000, generator=gp)
        grads.append(gr.numpy())
        par.grad = gr.clone()
        opt.step()
        mom = 1 - ((1 - 0.002) * (1 - np.exp(-step / 2000)))
        ema.mul_(1 - mom).add_(par.data, alpha=mom)
    tr["adam_w0"] = w.numpy(); tr["adam_grads"] = np.stack(grads); tr["adam_w3"] = par.data.numpy(); tr["ema3"] = ema.numpy()
    np.savez_compressed(os.path.join(HERE, "train.npz"), **tr)

    with open(os.path.join(HERE, "meta.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    for fn in sorted(os.listdir(HERE)):
        print(fn, os.path.getsize(os.path.join(HERE, fn)))


if __name__ == "__main__":
    main()
