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


def _parse_cpulist(text: str):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.update(range(int(a), int(b or a) + 1))
    return cpus


def bind_to_gpu_numa_node(device_index: int) -> dict:
    """Pin the calling process to the CPUs of the NUMA node the GPU hangs off (sysfs: the PCI device's ``numa_node`` and that
    node's ``cpulist``), so that staging buffers allocated afterwards (``pin_memory`` first-touches its pages in the calling
    thread) are node-local and the H2D DMA does not cross the inter-socket link.  One process per GPU (torchrun): call it before
    allocating pinned memory.  Returns what it did; never raises (containers may forbid it) - the copy path works unbound."""
    import os
    info = {"bound": False}
    try:
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        info["pci"] = bdf
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node < 0:
            info["why"] = "the platform reports no NUMA node for this device"
            return info
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = _parse_cpulist(f.read())
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        info["node_cpus"], info["allowed_cpus"] = len(cpus), len(allowed)
        if not use:
            info["why"] = "none of the node's CPUs is in this process's cpuset"
            return info
        os.sched_setaffinity(0, use)
        info["bound"], info["cpus"] = True, len(use)
    except Exception as e:  # noqa: BLE001
        info["why"] = f"{type(e).__name__}: {e}"
    return info


def h2d_ceiling_gbs(nbytes_per_copy: int, copies: int, device, reps: int = 3) -> float:
    """Raw pinned-host -> device copy rate (GB/s) for the same transfer pattern ``run_host_batch`` issues: ``copies`` back-to-back
    ``cudaMemcpyAsync`` of ``nbytes_per_copy`` each on one stream, timed with CUDA events (best of ``reps``).  The ceiling the
    end-to-end number is quoted against."""
    src = torch.empty(nbytes_per_copy, dtype=torch.uint8, pin_memory=True)
    src.zero_()
    dst = [torch.empty(nbytes_per_copy, dtype=torch.uint8, device=device) for _ in range(2)]
    st = _copy_stream(device)
    best = 0.0
    with torch.cuda.stream(st):
        dst[0].copy_(src, non_blocking=True)
        st.synchronize()
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            for i in range(copies):
                dst[i & 1].copy_(src, non_blocking=True)
            b.record(st)
            st.synchronize()
            best = max(best, nbytes_per_copy * copies / (a.elapsed_time(b) / 1e3) / 1e9)
    return best


def h2d_ceiling_tensor_gbs(x_host: torch.Tensor, chunk: int, device, reps: int = 2) -> float:
    """Raw pinned-host -> device copy rate (GB/s) of THE batch ``run_host_batch`` gets: the same host tensor, the same chunks,
    back to back on the copy stream into two alternating device buffers, nothing else running (best of ``reps``).  Unlike
    ``h2d_ceiling_gbs`` (one chunk-sized source buffer copied over and over, which the host's last-level cache can serve when
    several GPUs pull at once) every byte comes from a different place in host memory, as in the pipeline."""
    N = x_host.shape[0]
    chunk = max(1, min(int(chunk), N))
    dst = [torch.empty((chunk,) + tuple(x_host.shape[1:]), device=device, dtype=x_host.dtype) for _ in range(2)]
    st = _copy_stream(device)
    nbytes = x_host.numel() * x_host.element_size()
    best = 0.0
    with torch.cuda.stream(st):
        dst[0][:1].copy_(x_host[:1], non_blocking=True)
        st.synchronize()
        for _ in range(reps):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st)
            for i, s0 in enumerate(range(0, N, chunk)):
                n = min(chunk, N - s0)
                dst[i & 1][:n].copy_(x_host[s0:s0 + n], non_blocking=True)
            b.record(st)
            st.synchronize()
            best = max(best, nbytes / (a.elapsed_time(b) / 1e3) / 1e9)
    return best


_COPY_STREAMS = {}


def _copy_stream(dev) -> torch.cuda.Stream:
    key = (dev.type, dev.index)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(dev)
    return _COPY_STREAMS[key]
