"""Golden fixture for the batch builder, made by the LIVE reference's ``AudioDataset.__getitem__`` and
``collate_fn`` (dataset.py:103-164,276-283) in the dev container:

    python tests/golden/make_golden_collate.py

The reference reads audio with ``torchaudio.load`` (no decoder in this image): the call is replaced by a stub that returns the
seeded synthetic waveform of the requested (offset, length) window, everything after it is the reference's own code."""
import os
import sys
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
import dataset as ref_ds  # noqa: E402
os.chdir(ROOT)

import synth  # noqa: E402

SR, DUR = synth.COLLATE_SR, synth.COLLATE_DUR
collate_cases, waveform = synth.collate_cases, synth.collate_waveform


def main():
    cases = collate_cases()
    chan = {c[0]: c[1] for c in cases}

    def fake_load(filepath, frame_offset=0, num_frames=-1, backend=None):
        name = os.path.splitext(os.path.basename(filepath))[0]
        return waveform(name, chan[name], frame_offset, num_frames), SR
    ref_ds.torchaudio.load = fake_load
    ds = ref_ds.AudioDataset.__new__(ref_ds.AudioDataset)
    ds.audios_path, ds.extension, ds.sample_rate, ds.sample_duration, ds.ignore_index = "/nonexistent", "wav", SR, DUR, -100
    ds.class2idx = {"speech": 0, "music": 1}
    ds._samples = []
    for name, _, segs, gmm in cases:
        s = {"filename": name, "sample": np.array([[a, b, lab] for a, b, lab in segs], dtype=object)}
        if gmm is not None:
            s["group_minmax"] = gmm
        ds._samples.append(s)
    batch = [ds[i] for i in range(len(cases))]
    audio, targets = ref_ds.AudioDataset.collate_fn(batch)
    np.savez_compressed(os.path.join(HERE, "collate.npz"), audio=audio.numpy(), targets=targets.numpy())
    print(audio.shape, targets)


if __name__ == "__main__":
    main()
