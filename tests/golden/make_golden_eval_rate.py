"""Golden fixture for the chunker's FILE-RATE branch (inference.py:152-159: a file whose sample rate differs from the model's
``sample_rate`` goes through an extra torchaudio Resample before the network), made by the LIVE reference in the dev container:

    python tests/golden/make_golden_eval_rate.py      -> eval_rate16k.npz, eval_rate16k_results.csv

``torchaudio.load`` is stubbed to serve windows of a seeded synthetic 150-second waveform at 16 kHz; the model runs at 22.05 kHz.
Also stores torchaudio's own resampling of a short seeded signal for three rate pairs (the kernel-level fixture)."""
import os
import sys
import tempfile
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, HERE)
mp = types.ModuleType("matplotlib"); mp.pyplot = types.ModuleType("matplotlib.pyplot")
sys.modules["matplotlib"] = mp; sys.modules["matplotlib.pyplot"] = mp.pyplot
sys.path.insert(0, REF)
os.chdir(REF)
from modules import AudioDetectionNetwork  # noqa: E402
import inference as ref_inf  # noqa: E402
import torchaudio  # noqa: E402
os.chdir(ROOT)
import synth  # noqa: E402

SKIP = {"resampler.kernel", "melspectogram_tfmr.spectrogram.window", "melspectogram_tfmr.mel_scale.fb", "mfcc_tfmr.dct_mat",
        "mfcc_tfmr.MelSpectrogram.spectrogram.window", "mfcc_tfmr.MelSpectrogram.mel_scale.fb", "sm_anchors", "md_anchors",
        "lg_anchors", "taper_window"}


def main():
    torch.set_grad_enabled(False)
    m = AudioDetectionNetwork(2, config=f"{REF}/config/config.yaml")
    layout = {k: list(v.shape) for k, v in m.state_dict().items() if k not in SKIP}
    full = dict(m.state_dict()); full.update(synth.synth_state_dict(layout, seed=42))
    m.load_state_dict(full)
    m.eval()
    wav = synth.eval_waveform_rate(synth.EVAL_RATE2)

    def fake_load(filepath, frame_offset=0, num_frames=-1, backend=None):
        return wav[None, frame_offset:frame_offset + num_frames].clone(), synth.EVAL_RATE2
    ref_inf.torchaudio.load = fake_load
    captured = []
    orig_pmo = ref_inf.process_model_outputs

    def spy(*a, **k):
        seg, bidx = orig_pmo(*a, **k)
        captured.append((seg.clone(), bidx.clone()))
        return seg, bidx
    ref_inf.process_model_outputs = spy
    out_dir = tempfile.mkdtemp()
    ref_inf.evaluate_audio(m, "clips/long16k.wav", out_dir, input_sample_rate=synth.EVAL_SR, sample_duration=60, batch_size=2,
                           idx2class_map={0: "speech", 1: "music"}, device="cpu", iou_threshold=synth.EVAL_IOU,
                           conf_threshold=synth.EVAL_CONF)
    csvs = [os.path.join(dp, f) for dp, _, fs in os.walk(out_dir) for f in fs if f.endswith(".csv")]
    assert len(csvs) == 1, csvs
    text = open(csvs[0]).read()
    with open(os.path.join(HERE, "eval_rate16k_results.csv"), "w") as f:
        f.write(text)
    out = {"n_batches": np.int64(len(captured))}
    for i, c in enumerate(captured):
        out[f"seg{i}"], out[f"bidx{i}"] = c[0].numpy(), c[1].numpy()
    # kernel-level fixture: torchaudio.transforms.Resample on a short seeded signal
    g = torch.Generator().manual_seed(31)
    x = torch.randn(2, 1, 4001, generator=g)
    out["rs_x"] = x.numpy()
    for o, n in ((16000, 22050), (48000, 22050), (44100, 22050)):
        out[f"rs_{o}_{n}"] = torchaudio.transforms.Resample(orig_freq=o, new_freq=n)(x).numpy()
    np.savez_compressed(os.path.join(HERE, "eval_rate16k.npz"), **out)
    print(text, [tuple(c[0].shape) for c in captured], [c[1].tolist() for c in captured])


if __name__ == "__main__":
    main()
