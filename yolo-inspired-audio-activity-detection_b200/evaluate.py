"""The chunker around the hot path: ``inference.evaluate_audio`` (reference: inference.py:112-187) from the decoded waveform
on - 60-second windows, zero padding of the tail, batches of ``batch_size`` clips through the host-buffer pipeline
(``run_host_batch``: chunked H2D copies overlapped with the forward), segment times shifted to file time, run-length
merging of consecutive segments of one class.  File decoding and the CSV writer stay with the caller."""
from __future__ import annotations

from datetime import timedelta
from typing import Dict, List, Tuple

import ctypes as C
import math

import numpy as np
import torch

from . import _lib
from .frontend_consts import sinc_resample_kernel
from .hostpipe import run_host_batch
from .postprocess import process_model_outputs

_RS_CACHE: Dict[Tuple[int, int, str], Tuple[torch.Tensor, int, int, int]] = {}


def resample_to_model_rate(x: torch.Tensor, orig_rate: int, new_rate: int) -> torch.Tensor:
    """``torchaudio.transforms.Resample(orig_freq, new_freq)`` (defaults: Hann-windowed sinc, lowpass_filter_width 6, rolloff
    0.99) of a CUDA batch ``x [B, 1, L]`` (fp32 or int16 PCM) -> fp32 ``[B, 1, ceil(new * L / orig)]`` on the sm_100a kernel
    ``yad_resample_sinc`` (reference call site: inference.py:152-159)."""
    if not x.is_cuda:
        raise RuntimeError("resample_to_model_rate needs a CUDA tensor (no CPU fallback)")
    if x.dtype not in (torch.float32, torch.int16):
        raise ValueError("resample_to_model_rate takes fp32 or int16 PCM")
    dev = x.device
    key = (int(orig_rate), int(new_rate), str(dev))
    if key not in _RS_CACHE:
        k, width, o, n = sinc_resample_kernel(int(orig_rate), int(new_rate))
        _RS_CACHE[key] = (k[:, 0, :].contiguous().to(dev), width, o, n)
    k, width, o, n = _RS_CACHE[key]
    B, _, L = x.shape
    Lout = int(math.ceil(n * L / o))
    x = x.contiguous()
    out = torch.empty((B, 1, Lout), device=dev, dtype=torch.float32)
    lib = _lib.init(dev.index if dev.index is not None else torch.cuda.current_device())
    with torch.cuda.device(dev):
        rc = lib.yad_resample_sinc(x.data_ptr(), 1 if x.dtype == torch.int16 else 0, B, L, o, n, width, k.data_ptr(), out.data_ptr(),
                                   Lout, C.c_void_p(torch.cuda.current_stream(dev).cuda_stream))
    _lib.check(rc, "resample_sinc")
    return out


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
    file_rate = int(og_sample_rate) != in_rate      # the extra Resample of inference.py:152-159 (yad_resample_sinc)
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
        if file_rate:
            dev = next(model.parameters()).device
            with torch.no_grad():
                xr = resample_to_model_rate(xp.to(dev, non_blocking=True), int(og_sample_rate), in_rate)
                out = model(xr, combine_scales=True)
            segments, batch_idxs = process_model_outputs(out, iou_threshold, conf_threshold, sample_duration, True)   # raises if empty
            segments, batch_idxs = segments.cpu(), batch_idxs.cpu()
        else:
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
