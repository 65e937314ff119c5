"""``process_model_outputs`` drop-in (reference: inference.py:42-110) on the sm_100a NMS kernels."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib


def nms_raw(outputs: torch.Tensor, iou_threshold: float, conf_threshold: float, sample_duration: float = 60,
            return_start_end: bool = True, _h: float = 10, want_taps: bool = False, want_keep: bool = True) -> Dict[str, torch.Tensor]:
    """Runs the per-clip NMS kernel and returns its raw device outputs:
    keep [B,P] i32 (descending score, -1 padded), n_keep [B], seg_rows [B,P,5], n_seg [B] (+ conf/boxes taps).
    ``want_keep=False`` (what ``process_model_outputs`` needs: the reference never returns the keep list) lets the kernel stop
    its greedy scan at the confidence threshold."""
    if not outputs.is_cuda:
        raise RuntimeError("yad_b200.process_model_outputs needs a CUDA tensor (no CPU fallback)")
    if outputs.ndim != 3:
        outputs = outputs.unsqueeze(0)
    assert outputs.ndim == 3, "input is expected to have 2 or 3 dimensions"
    dev = outputs.device
    lib = _lib.init(dev.index if dev.index is not None else torch.cuda.current_device())
    x = outputs.contiguous().float()
    B, P, E = x.shape
    nc = E - 3
    r = {
        "seg_rows": torch.empty((B, P, 5), device=dev, dtype=torch.float32),
        "n_seg": torch.empty((B,), device=dev, dtype=torch.int32),
    }
    if want_keep:
        r["keep"] = torch.empty((B, P), device=dev, dtype=torch.int32)
        r["n_keep"] = torch.empty((B,), device=dev, dtype=torch.int32)
    if want_taps:
        r["conf"] = torch.empty((B, P), device=dev, dtype=torch.float32)
        r["boxes"] = torch.empty((B, P, 2), device=dev, dtype=torch.float32)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = lib.yad_nms(x.data_ptr(), B, P, nc, float(iou_threshold), float(conf_threshold), float(sample_duration), float(_h),
                         1 if return_start_end else 0, _lib.ptr(r.get("keep")), _lib.ptr(r.get("n_keep")), _lib.ptr(r.get("conf")),
                         _lib.ptr(r.get("boxes")), r["seg_rows"].data_ptr(), r["n_seg"].data_ptr(), stream)
    _lib.check(rc, "nms")
    return r


def segments_device(outputs: torch.Tensor, iou_threshold: float = 0.05, conf_threshold: float = 0.5, sample_duration: float = 60,
                    return_start_end: bool = True, _h: int = 10) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """The device side of ``process_model_outputs`` without its host synchronisation: per-clip NMS + compaction.  Returns
    (segments [B*P,5], batch_idxs [B*P], total [1] i64) on the device; rows [0, total) are valid."""
    r = nms_raw(outputs, iou_threshold, conf_threshold, sample_duration, return_start_end, _h, want_keep=False)
    dev = r["seg_rows"].device
    B, P = r["seg_rows"].shape[:2]
    lib = _lib.load()
    segments = torch.empty((B * P, 5), device=dev, dtype=torch.float32)
    batch_idxs = torch.empty((B * P,), device=dev, dtype=torch.int64)
    total = torch.empty((1,), device=dev, dtype=torch.int64)
    with torch.cuda.device(dev):
        stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        rc = lib.yad_compact_segments(r["seg_rows"].data_ptr(), r["n_seg"].data_ptr(), B, P, segments.data_ptr(),
                                      batch_idxs.data_ptr(), total.data_ptr(), stream)
    _lib.check(rc, "compact_segments")
    return segments, batch_idxs, total


def process_model_outputs(outputs: torch.Tensor, iou_threshold: float = 0.05, conf_threshold: float = 0.5,
                          sample_duration: float = 60, return_start_end: bool = True, _h: int = 10,
                          ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Same signature and return as the reference: (segments [K,5] f32 = [conf, obj_logit, label, start, end],
    batch_idxs [K] i64), clips in order, segments sorted by centre inside a clip.

    Class-agnostic per-clip NMS on un-offset coordinates (== torchvision.batched_nms whenever it takes its
    per-index loop, and at B = 1).  Raises ValueError when nothing passes ``conf_threshold`` - the
    reference fails the same way (torch.cat of an empty list, inference.py:100)."""
    segments, batch_idxs, total = segments_device(outputs, iou_threshold, conf_threshold, sample_duration, return_start_end, _h)
    K = int(total.item())            # the output shape is data dependent: one host sync, like the reference
    if K == 0:
        raise ValueError("no segment passed conf_threshold (the reference raises here too: torch.cat of an empty list)")
    return segments[:K], batch_idxs[:K]
