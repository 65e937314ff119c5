"""Deterministic synthetic inputs and weights shared by the golden generator, the tests and bench.py.

Everything is drawn from CPU ``torch.Generator`` streams, so the same tensors are reproduced on any
machine with the same torch version (the golden outputs were produced by the live reference on them)."""
from __future__ import annotations

import hashlib
import math
from typing import Dict, Sequence

import torch


def _seed_of(name: str, seed: int) -> int:
    return int.from_bytes(hashlib.sha256(f"{seed}:{name}".encode()).digest()[:6], "little")


def synth_state_dict(layout: Dict[str, Sequence[int]], seed: int = 42) -> Dict[str, torch.Tensor]:
    """Weights for every key of a (train-form) state-dict layout: xavier-like conv weights, randomised
    BatchNorm statistics (so that BN folding bugs cannot hide), small biases.  Frontend buffers and anchors
    are NOT produced here (they are constants of the config)."""
    out = {}
    for k, shp in layout.items():
        shp = tuple(shp)
        g = torch.Generator().manual_seed(_seed_of(k, seed))
        if k.endswith("num_batches_tracked"):
            out[k] = torch.tensor(0, dtype=torch.int64)
        elif k.endswith("running_mean"):
            out[k] = torch.randn(shp, generator=g) * 0.5
        elif k.endswith("running_var"):
            out[k] = torch.rand(shp, generator=g) * 1.5 + 0.5
        elif len(shp) == 4:  # conv weight
            fan_in, fan_out = shp[1] * shp[2] * shp[3], shp[0] * shp[2] * shp[3]
            bound = math.sqrt(6.0 / (fan_in + fan_out))
            out[k] = (torch.rand(shp, generator=g) * 2 - 1) * bound
        elif ".norm.weight" in k or ".bn" in k and k.endswith(".weight") or k.endswith("identity.weight") \
                or ".downsample.1.weight" in k:
            out[k] = torch.rand(shp, generator=g) + 0.5
        elif k.endswith(".weight") and len(shp) == 1:   # any other BatchNorm weight (custom backbone: first_conv.1, _layer.1/.4)
            out[k] = torch.rand(shp, generator=g) + 0.5
        elif k.endswith(".bias"):
            out[k] = torch.randn(shp, generator=g) * 0.2
        else:
            raise KeyError(f"synth_state_dict: no rule for {k} {shp}")
    return out


def synth_clips(B: int, L: int, seed: int = 1000, sample_rate: int = 22050, silence_tail_every: int = 4) -> torch.Tensor:
    """[B,1,L] f32: 0.1*noise + 3..6 gated tone bursts; every ``silence_tail_every``-th clip ends in digital
    silence (mirrors the reference's zero padding of short clips, dataset.py:150-155)."""
    x = torch.empty(B, 1, L)
    t = torch.arange(L, dtype=torch.float32) / sample_rate
    dur = L / sample_rate
    for i in range(B):
        g = torch.Generator().manual_seed(seed + i)
        s = 0.1 * torch.randn(L, generator=g)
        nb = int(torch.randint(3, 7, (1,), generator=g))
        for _ in range(nb):
            f = float(torch.rand(1, generator=g)) * 6900.0 + 100.0
            t0 = float(torch.rand(1, generator=g)) * dur * 0.8
            d = float(torch.rand(1, generator=g)) * dur * 0.2 + dur * 0.02
            m = (t >= t0) & (t < t0 + d)
            s = s + 0.5 * torch.sin(2 * math.pi * f * t) * m
        if silence_tail_every and i % silence_tail_every == silence_tail_every - 1:
            cut = int((0.35 + 0.55 * float(torch.rand(1, generator=g))) * L)
            s[cut:] = 0.0
        x[i, 0] = s
    return x


def synth_heads(B: int, P: int = 630, nc: int = 2, seed: int = 7, adversarial: bool = True) -> torch.Tensor:
    """Decoded head tensor [B,P,3+nc] for NMS tests: logits ~ N(0,2), centres ~ U(0,60), widths ~ U(0,30),
    with exact duplicates, zero widths and score ties planted on purpose."""
    g = torch.Generator().manual_seed(seed)
    o = torch.empty(B, P, 3 + nc)
    o[..., : 1 + nc] = torch.randn(B, P, 1 + nc, generator=g) * 2.0
    o[..., -2] = torch.rand(B, P, generator=g) * 60.0
    o[..., -1] = torch.rand(B, P, generator=g) * 30.0
    if adversarial:
        for b in range(B):
            o[b, 10] = o[b, 3]                      # exact duplicate row (score tie, IoU 1)
            o[b, 50, : 1 + nc] = o[b, 40, : 1 + nc]  # score tie, different box
            o[b, 70, -1] = 0.0                      # zero-width boxes (NaN IoU between them)
            o[b, 71, -1] = 0.0
            o[b, 71, -2] = o[b, 70, -2]
            o[b, 90, -2], o[b, 90, -1] = 0.0, 5.0   # clipped at 0
            o[b, 91, -2], o[b, 91, -1] = 60.0, 9.0  # clipped at 60
    return o


def synth_targets(B: int, seed: int = 11, duration: float = 60.0) -> torch.Tensor:
    """[T,4] = (batch_idx, cls, centre_s, dur_s): per clip 5-7 contiguous segments tiling [0, duration];
    the last clip carries one ignore-label (-100) target like a zero-padded clip (dataset.py:156-160)."""
    g = torch.Generator().manual_seed(seed)
    rows = []
    for b in range(B):
        n = int(torch.randint(5, 8, (1,), generator=g))
        cuts = torch.sort(torch.rand(n - 1, generator=g) * duration).values
        edges = torch.cat([torch.zeros(1), cuts, torch.full((1,), duration)])
        for i in range(n):
            s, e = float(edges[i]), float(edges[i + 1])
            cls = int(torch.randint(0, 2, (1,), generator=g))
            if b == B - 1 and i == n - 1:
                cls = -100
            rows.append([b, cls, 0.5 * (s + e), e - s])
    return torch.tensor(rows, dtype=torch.float32)


# ---- batch-builder cases (tests/golden/make_golden_collate.py)
COLLATE_SR, COLLATE_DUR = 2000, 6          # a small sample rate keeps the fixture tiny; the code path does not depend on it


def collate_cases():
    """(filename, channels, segments [(start, end, label)], group_minmax or None)"""
    return [
        ("a", 1, [(0.0, 1.25, "speech"), (1.25, 3.5, "music"), (3.5, 6.0, "speech")], None),            # full length
        ("b", 2, [(0.0, 2.0, "music"), (2.0, 4.1, "speech")], None),                                    # stereo, short -> pad label
        ("c", 1, [(12.0, 13.5, "speech"), (13.5, 15.75, "music"), (15.75, 17.0, "speech")], (12.0, 18.0)),  # grouped, short
        ("d", 3, [(6.0, 9.0, "music"), (9.0, 12.0, "speech")], (6.0, 12.0)),                            # 3 channels, grouped, full
    ]


def collate_waveform(name: str, channels: int, frame_offset: int, num_frames: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(sum(map(ord, name)) * 7919 + frame_offset)
    return (torch.randn(channels, num_frames, generator=g) * 0.25).clamp(-1, 1)


# ---- file-level chunker case (tests/golden/make_golden_eval.py)
EVAL_SR, EVAL_IOU, EVAL_CONF = 22050, 0.1, 0.2


def eval_waveform() -> torch.Tensor:
    """270 s mono waveform (4.5 clips of 60 s: three batches of 2, 2 and 1 zero-padded clip)."""
    x = synth_clips(5, 22050 * 60, seed=4000, silence_tail_every=0)
    return x.reshape(-1)[: int(270 * 22050)].contiguous()


EVAL_RATE2 = 16000


def eval_waveform_rate(rate: int = EVAL_RATE2) -> torch.Tensor:
    """150 s mono waveform at another file rate (2.5 clips of 60 s: two batches of 2 and 1 zero-padded clip): the chunker's
    file-rate branch (inference.py:152-159)."""
    x = synth_clips(3, rate * 60, seed=4100, silence_tail_every=0)
    return x.reshape(-1)[: int(150 * rate)].contiguous()
