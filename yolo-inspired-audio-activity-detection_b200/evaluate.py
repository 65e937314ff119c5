"""The chunker around the hot path: ``inference.evaluate_audio`` (reference: inference.py:112-187) from the decoded waveform
on - 60-second windows, zero padding of the tail, batches of ``batch_size`` clips through the host-buffer pipeline
(``run_host_batch``: chunked H2D copies overlapped with the forward), segment times shifted to file time, run-length
merging of consecutive segments of one class.  File decoding and the CSV writer stay with the caller."""
from __future__ import annotations

from datetime import timedelta
from typing import Dict, List, Tuple

import numpy as np
import torch

from .hostpipe import run_host_batch


def evaluate_waveform(model, waveform: torch.Tensor, og_sample_rate: int, sample_duration: float, batch_size: int,
                      idx2class_map: Dict[int, str], iou_threshold: float = 0.1, conf_threshold: float = 0.65, chunk: int = 32,
                      ) -> Tuple[torch.Tensor, torch.Tensor, List[dict]]:
    """waveform: host tensor [n] or [C, n], fp32 or int16 PCM, at ``og_sample_rate``.  Returns (segments [K,5] with start / end in
    file time, batch_idxs [K], rle rows [{"start": timedelta, "end": timedelta, "class": str}]) - the rows the reference writes
    to ``<file>_results.csv``.

    Faithful to the reference, including two quirks: the clip index offset of a later batch is the LAST clip index seen so far
    (``batch_idxs += batch_idxs_list[-1][-1]``, inference.py:176-177), not the clip count; a batch without any segment above
    ``conf_threshold`` raises ``ValueError`` (the reference's ``torch.cat`` of an empty list)."""
    in_rate = int(model.config["sample_rate"])
    if int(og_sample_rate) != in_rate:
        raise NotImplementedError(f"yad_b200.evaluate_waveform: the file is at {og_sample_rate} Hz, the model expects {in_rate} Hz; "
                                  "the extra torchaudio Resample of inference.py:152-159 is not built")
    if waveform.is_cuda:
        raise ValueError("evaluate_waveform takes the decoded waveform on the host")
    wav = waveform if waveform.ndim == 2 else waveform.unsqueeze(0)
    if wav.dtype not in (torch.float32, torch.int16):
        raise ValueError("evaluate_waveform takes fp32 or int16 PCM")
    sample_size = int(sample_duration * og_sample_rate)
    batch_start, batch_end = 0, batch_size * sample_duration
    segments_list, batch_idxs_list = [], []
    while True:
        lo, num = int(batch_start * og_sample_rate), int((batch_end - batch_start) * og_sample_rate)
        x = wav[:, lo:lo + num]
        if x.shape[-1] == 0:
            break
        x = x.squeeze(0)
        if x.ndim == 2 and x.shape[0] > 1:
            if x.dtype == torch.int16:
                raise NotImplementedError("multi-channel int16 input: mix down to mono (or pass fp32) first")
            x = x.mean(dim=0).squeeze(0)
        if x.shape[0] % sample_size != 0:
            nbatch = int(np.ceil(x.shape[0] / sample_size))
            x = torch.cat([x, torch.zeros((nbatch * sample_size - x.shape[0],), dtype=x.dtype)], dim=0)
        x = x.reshape(-1, 1, sample_size)
        if len(idx2class_map) < model.num_classes:
            raise RuntimeError("model output does not match idx2class mapping")
        xp = x if x.is_pinned() else x.contiguous().pin_memory()
        segments, batch_idxs = run_host_batch(model, xp, iou_threshold, conf_threshold, chunk=chunk, sample_duration=sample_duration,
                                              return_start_end=True)
        if segments is None:
            raise ValueError("no segment passed conf_threshold in this batch (the reference raises here too: torch.cat of an empty list)")
        if len(batch_idxs_list) > 0:
            batch_idxs = batch_idxs + batch_idxs_list[-1][-1]
        segments_list.append(segments)
        batch_idxs_list.append(batch_idxs)
        batch_start = batch_end
        batch_end += batch_size * sample_duration
    segments = torch.cat(segments_list, dim=0)
    batch_idxs = torch.cat(batch_idxs_list, dim=0)
    segments[..., -2:] = segments[..., -2:] + (batch_idxs.unsqueeze(-1) * sample_duration)
    return segments, batch_idxs, rle_rows(segments, idx2class_map)


def rle_rows(segments: torch.Tensor, idx2class_map: Dict[int, str]) -> List[dict]:
    """Run-length merge of consecutive segments with the same class (inference.py:189-198)."""
    rows: List[dict] = []
    for i in range(segments.shape[0]):
        s = segments[i]
        start = timedelta(seconds=round(s[-2].item(), 2))
        end = timedelta(seconds=round(s[-1].item(), 2))
        cls = idx2class_map[int(s[2].item())]
        if len(rows) == 0 or rows[-1]["class"] != cls:
            rows.append({"start": start, "end": end, "class": cls})
            continue
        rows[-1]["end"] = end
    return rows
