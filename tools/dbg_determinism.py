import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import torch, numpy as np
import yad_b200, synth
from oracle import ref_port as O
import train_helpers as TH
meta = json.load(open(os.path.join(ROOT, "tests/golden/meta.json")))
SKIP = {"resampler.kernel", "melspectogram_tfmr.spectrogram.window", "melspectogram_tfmr.mel_scale.fb", "mfcc_tfmr.dct_mat",
        "mfcc_tfmr.MelSpectrogram.spectrogram.window", "mfcc_tfmr.MelSpectrogram.mel_scale.fb", "sm_anchors", "md_anchors", "lg_anchors", "taper_window"}
layout = {k: v for k, v in meta["layout_train"].items() if k not in SKIP}
sd = synth.synth_state_dict(layout, seed=42); sd.update(O.frontend_constants())
for k in ("sm_anchors", "md_anchors", "lg_anchors"):
    sd[k] = torch.tensor(meta["anchors"][k], dtype=torch.float32)
sd["taper_window"] = torch.empty(0)
dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
x = synth.synth_clips(B, TH.TRAIN_L, seed=3000).to(dev); tg = synth.synth_targets(B, seed=21, duration=16.0).to(dev)
cfg = yad_b200.default_config(); cfg["dropout"] = 0.0
lf = yad_b200.AudioDetectionLoss(cfg["anchors"], 2, sample_duration=60, **cfg["train_config"]["loss_config"])
rel = lambda a, b: float((a.double() - b.double()).norm() / b.double().norm())
for mode in ("f32", "tf32"):
    m = yad_b200.AudioDetectionNetwork(2, config=cfg, train_dtype=mode); m.load_state_dict(sd); m = m.to(dev).train()
    runs = []
    for it in range(3):
        m.load_state_dict(sd)
        for p in m.parameters(): p.grad = None
        with torch.enable_grad():
            preds = m(x); loss, _ = lf(preds, tg); loss.backward()
        g = torch.cat([p.grad.reshape(-1) for p in m.parameters()])
        runs.append((float(loss), torch.cat([p.detach().reshape(-1) for p in preds]).clone(), g.clone()))
    print(mode, "B", B, "loss", [r[0] for r in runs], "pred maxdiff run0-1", float((runs[0][1] - runs[1][1]).abs().max()),
          "grad rel run0-1", rel(runs[0][2], runs[1][2]), "run1-2", rel(runs[1][2], runs[2][2]))

# ---- which TF32 forward conv is not reproducible?
m = yad_b200.AudioDetectionNetwork(2, config=cfg, train_dtype="tf32"); m.load_state_dict(sd); m = m.to(dev).train()
eng = m._train_engine()
orig = eng._conv_tf32
logs = []
def spy(x, conv, out, need_dx):
    o = orig(x, conv, out, need_dx)
    logs[-1].append((tuple(conv.weight.shape), tuple(conv.stride), (x.B, x.H, x.W), o.buf.clone()))
    return o
eng._conv_tf32 = spy
for it in range(2):
    m.load_state_dict(sd)
    logs.append([])
    with torch.no_grad():
        m(x)
for (w, s_, g, a), (_, _, _, b) in zip(logs[0], logs[1]):
    d = float((a - b).abs().max()); sc = float(a.abs().max())
    print("conv", w, s_, g, "maxdiff %.3e scale %.3e" % (d, sc), "NONDET" if d > 1e-5 * sc else "")
