"""Host-buffer inference pipeline: the accelerated counterpart of the batch loop in ``inference.evaluate_audio``
(reference: inference.py:126-174) for callers whose clips live in (pinned) host memory.

A 60 s clip is 5.3 MB of fp32 PCM, so a 512-clip batch is 2.7 GB over PCIe (~50 ms) against ~6 ms of GPU work:
end to end the path is bound by the host->device copy.  ``run_host_batch`` therefore cuts the batch into chunks and
double-buffers them: the copy of chunk i+1 (copy stream) overlaps the forward of chunk i (compute stream); the
predictions of all chunks are post-processed once at the end."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from .postprocess import process_model_outputs


def run_host_batch(model, x_host: torch.Tensor, iou_threshold: float = 0.05, conf_threshold: float = 0.5, chunk: int = 64,
                   sample_duration: float = 60, return_start_end: bool = True,
                   ) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """x_host [N,1,L] f32 (or int16 PCM, half the bytes over PCIe) on the host (pinned for an asynchronous copy) -> (segments [K,5], batch_idxs [K]) on the host, as
    ``process_model_outputs`` returns them (``(None, None)`` when nothing passes the confidence threshold).
    ``model`` is a ``yad_b200.AudioDetectionNetwork`` on a CUDA device."""
    if x_host.is_cuda:
        raise ValueError("run_host_batch takes a host tensor; call the model directly for device tensors")
    dev = next(model.parameters()).device
    N, _, L = x_host.shape
    chunk = max(1, min(int(chunk), N))
    compute = torch.cuda.current_stream(dev)
    copy = _copy_stream(dev)
    if x_host.dtype not in (torch.float32, torch.int16):
        raise ValueError("run_host_batch takes fp32 PCM or 16-bit PCM (int16; x / 32768 is applied on the GPU)")
    bufs = [torch.empty((chunk, 1, L), device=dev, dtype=x_host.dtype) for _ in range(2)]
    copied = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    preds = []
    with torch.no_grad():
        starts = list(range(0, N, chunk))
        copy.wait_stream(compute)
        for i, s in enumerate(starts):
            n, b = min(chunk, N - s), i & 1
            with torch.cuda.stream(copy):
                if i >= 2:
                    copy.wait_event(consumed[b])           # the forward that read this buffer two chunks ago is done
                bufs[b][:n].copy_(x_host[s:s + n], non_blocking=True)
                copied[b].record(copy)
            compute.wait_event(copied[b])
            preds.append(model(bufs[b][:n], combine_scales=True))
            consumed[b].record(compute)
        out = torch.cat(preds, 0) if len(preds) > 1 else preds[0]
        try:
            seg, bidx = process_model_outputs(out, iou_threshold, conf_threshold, sample_duration, return_start_end)
        except ValueError:
            return None, None
        return seg.cpu(), bidx.cpu()


_COPY_STREAMS = {}


def _copy_stream(dev) -> torch.cuda.Stream:
    key = (dev.type, dev.index)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(dev)
    return _COPY_STREAMS[key]
