"""Golden fixture for the file-level chunker, made by the LIVE reference's ``inference.evaluate_audio``
(inference.py:112-209) in the dev container:

    python tests/golden/make_golden_eval.py

``torchaudio.load`` (no decoder in this image) is replaced by a stub that serves windows of a seeded synthetic 270-second
waveform; the reference model carries the synthetic weights of synth.py.  The CSV the reference writes is the fixture."""
import os
import sys
import tempfile
import types

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
    wav = synth.eval_waveform()

    def fake_load(filepath, frame_offset=0, num_frames=-1, backend=None):
        return wav[None, frame_offset:frame_offset + num_frames].clone(), synth.EVAL_SR
    ref_inf.torchaudio.load = fake_load
    captured = []
    orig_pmo = ref_inf.process_model_outputs

    def spy(*a, **k):                      # record what the reference's own post-processing returned for every batch
        seg, bidx = orig_pmo(*a, **k)
        captured.append((seg.clone(), bidx.clone()))
        return seg, bidx
    ref_inf.process_model_outputs = spy
    out_dir = tempfile.mkdtemp()
    ref_inf.evaluate_audio(m, "clips/long.wav", out_dir, input_sample_rate=synth.EVAL_SR, sample_duration=60, batch_size=2,
                           idx2class_map={0: "speech", 1: "music"}, device="cpu", iou_threshold=synth.EVAL_IOU,
                           conf_threshold=synth.EVAL_CONF)
    csvs = [os.path.join(dp, f) for dp, _, fs in os.walk(out_dir) for f in fs if f.endswith(".csv")]
    assert len(csvs) == 1, csvs
    text = open(csvs[0]).read()
    with open(os.path.join(HERE, "eval_long_results.csv"), "w") as f:
        f.write(text)
    import numpy as np
    np.savez_compressed(os.path.join(HERE, "eval_long.npz"), n_batches=np.int64(len(captured)),
                        **{f"seg{i}": c[0].numpy() for i, c in enumerate(captured)},
                        **{f"bidx{i}": c[1].numpy() for i, c in enumerate(captured)})
    print(text, [tuple(c[0].shape) for c in captured], [c[1].tolist() for c in captured], [sorted(set(c[0][:, 2].tolist())) for c in captured])


if __name__ == "__main__":
    main()
