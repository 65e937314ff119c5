"""Diagnostic: per-parameter gradient error of the GPU train path and of the fp32 oracle, both against the oracle in fp64."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import json, torch, numpy as np
import yad_b200, synth
from oracle import ref_port as O
import train_helpers as TH
from yad_b200.train_engine import run_train_forward
torch.backends.cudnn.allow_tf32 = False
meta = json.load(open(os.path.join(ROOT, "tests/golden/meta.json")))
SKIP = {"resampler.kernel", "melspectogram_tfmr.spectrogram.window", "melspectogram_tfmr.mel_scale.fb", "mfcc_tfmr.dct_mat",
        "mfcc_tfmr.MelSpectrogram.spectrogram.window", "mfcc_tfmr.MelSpectrogram.mel_scale.fb", "sm_anchors", "md_anchors", "lg_anchors", "taper_window"}
layout = {k: v for k, v in meta["layout_train"].items() if k not in SKIP}
sd = synth.synth_state_dict(layout, seed=42); sd.update(O.frontend_constants())
for k in ("sm_anchors", "md_anchors", "lg_anchors"):
    sd[k] = torch.tensor(meta["anchors"][k], dtype=torch.float32)
sd["taper_window"] = torch.empty(0)
x, tg = TH.train_inputs()
xs32 = O.frontend(x, sd)["x_spectral"]

def oracle_from_xs(sd, xs, dtype):
    sd = {k: (v.clone().to(dtype) if v.dtype.is_floating_point else v.clone()) for k, v in sd.items()}
    names = TH.param_names(sd)
    with torch.enable_grad():
        for k in names: sd[k].requires_grad_(True)
        O._BN_TRAINING[0] = True
        fm = O.backbone(sd, xs.to(dtype), (2, 2, 2, 2)); heads = O.neck(sd, fm)
        O._BN_TRAINING[0] = False
        preds = O.decode(heads, sd, 256146, 256, 2, TH.TRAIN_CFG, combine_scales=False)
        loss, _ = O.detection_loss(preds, tg.to(dtype), O.DEFAULT_CONFIG["anchors"], 2)
        loss.backward()
    return loss.item(), {k: sd[k].grad for k in names}
l64, g64 = oracle_from_xs(sd, xs32, torch.float64)
l32, g32 = oracle_from_xs(sd, xs32, torch.float32)
dev = torch.device("cuda", 0)
cfg = yad_b200.default_config(); cfg["dropout"] = 0.0
m = yad_b200.AudioDetectionNetwork(2, config=cfg, train_dtype=(sys.argv[1] if len(sys.argv) > 1 else "f32")); m.load_state_dict(sd); m = m.to(dev).train()
lf = yad_b200.AudioDetectionLoss(cfg["anchors"], 2, sample_duration=60, **cfg["train_config"]["loss_config"])
with torch.enable_grad():
    preds = run_train_forward(m, m._train_engine(), xs32.to(dev).contiguous(), 256, 256146)
    loss, _ = lf(preds, tg.to(dev)); loss.backward()
print("loss f64 %.8f o32 %.8f gpu %.8f" % (l64, l32, float(loss)))
rel = lambda a, b: float((a.double().cpu().reshape(-1) - b.double().reshape(-1)).norm() / max(float(b.double().norm()), 1e-30))
for k, p in m.named_parameters():
    print(f"{k:70s} |g| {float(g64[k].norm()):.3e}  gpu-f64 {rel(p.grad, g64[k]):.2e}  o32-f64 {rel(g32[k], g64[k]):.2e}")

# ---- yardstick: the ORACLE itself on the GPU with cuDNN's TF32 convolutions (torch default) vs fp32
def oracle_cuda(allow):
    torch.backends.cudnn.allow_tf32 = allow
    sdc = {k: v.to(dev) for k, v in sd.items()}
    names = TH.param_names(sdc)
    with torch.enable_grad():
        for k in names: sdc[k].requires_grad_(True)
        O._BN_TRAINING[0] = True
        heads = O.neck(sdc, O.backbone(sdc, xs32.to(dev), (2, 2, 2, 2)))
        O._BN_TRAINING[0] = False
        import types
        return heads, sdc, names
try:
    for allow in (False, True):
        heads, sdc, names = oracle_cuda(allow)
        # decode + loss on the CPU oracle (index ops); move heads to CPU keeping the graph
        preds = O.decode([h.cpu() for h in heads], {k: v.cpu() if not v.requires_grad else v.cpu() for k, v in sdc.items()}, 256146, 256, 2, TH.TRAIN_CFG, combine_scales=False)
        loss, _ = O.detection_loss(preds, tg, O.DEFAULT_CONFIG["anchors"], 2)
        loss.backward()
        errs = {k: rel(sdc[k].grad, g64[k]) for k in names if sdc[k].grad is not None and float(g64[k].norm()) > 1e-9}
        ks = ["feature_extractor.conv1.weight", "feature_extractor.layer2.0.conv1.weight", "feature_extractor.layer4.1.conv2.weight",
              "multiscale_module.cspsppf.conv7.conv.weight", "multiscale_module.rep_block4_1.blocks.0.conv1x1.norm.bias"]
        print("torch-cuda allow_tf32=%s loss %.6f" % (allow, float(loss)), {k.split("module.")[-1]: "%.2e" % errs[k] for k in ks}, "median %.2e max %.2e" % (sorted(errs.values())[len(errs) // 2], max(errs.values())))
except Exception as e:
    import traceback; traceback.print_exc()
