"""Host-side construction of the frontend's constant buffers and of the kernel-side packed forms.

The values equal the registered buffers of the reference's torchaudio transforms (state_dict names
in SURVEY App. F); formulas follow torchaudio 2.11 functional.py:1305-1400 (sinc kernel), :518-587
(mel filterbank), :636-668 (DCT-II ortho)."""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch

from ._lib import FE_QW


def sinc_resample_kernel(orig_freq: int, new_freq: int, lowpass_filter_width: int = 6,
                         rolloff: float = 0.99) -> Tuple[torch.Tensor, int, int, int]:
    g = math.gcd(int(orig_freq), int(new_freq))
    o, n = int(orig_freq) // g, int(new_freq) // g
    base = min(o, n) * rolloff
    width = math.ceil(lowpass_filter_width * o / base)
    idx = torch.arange(-width, width + o, dtype=torch.float64)[None, None] / o
    # the phase term is int64 / int -> float32 in torchaudio (dtype=None), then promoted to float64
    t = torch.arange(0, -n, -1)[:, None, None] / n + idx
    t = (t * base).clamp(-lowpass_filter_width, lowpass_filter_width)
    window = torch.cos(t * math.pi / lowpass_filter_width / 2) ** 2
    t = t * math.pi
    k = torch.where(t == 0, torch.tensor(1.0, dtype=torch.float64), t.sin() / t) * window * (base / o)
    return k.to(torch.float32), width, o, n


def mel_filterbank(n_freqs: int, f_min: float, f_max: float, n_mels: int, sample_rate: int, norm, mel_scale) -> torch.Tensor:
    if mel_scale != "htk" or norm != "slaney":
        raise NotImplementedError("only mel_scale='htk', norm='slaney' (the reference config) are built")
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = torch.max(torch.zeros(1), torch.min(down, up))
    enorm = 2.0 / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
    return fb * enorm.unsqueeze(0)


def dct_ortho(n_mfcc: int, n_mels: int) -> torch.Tensor:
    n = torch.arange(float(n_mels))
    k = torch.arange(float(n_mfcc)).unsqueeze(1)
    dct = torch.cos(math.pi / float(n_mels) * (n + 0.5) * k)
    dct[0] *= 1.0 / math.sqrt(2.0)
    dct *= math.sqrt(2.0 / float(n_mels))
    return dct.t().contiguous()


def pack_resample_taps(kernel: torch.Tensor) -> Dict[str, object]:
    """resampler.kernel [P,1,KW] -> per phase-quad taps over the quad's common window (kernel ABI)."""
    k = kernel.detach().to("cpu", torch.float32)[:, 0, :].numpy()
    P, KW = k.shape
    if P % 4:
        raise NotImplementedError(f"a number of resample phases that is not a multiple of 4 ({P}) is not supported by the kernel")
    taps = np.zeros((P // 4, 4, FE_QW), np.float32)
    base = np.zeros((P // 4,), np.int32)
    for u in range(P // 4):
        nz = np.nonzero((k[4 * u: 4 * u + 4] != 0).any(axis=0))[0]
        lo, hi = (int(nz[0]), int(nz[-1]) + 1) if nz.size else (0, 1)
        if hi - lo > FE_QW:
            raise NotImplementedError(
                f"resample phase quad {u} spans {hi - lo} taps > {FE_QW}: this sample-rate pair is not supported")
        base[u] = lo
        seg = k[4 * u: 4 * u + 4, lo:min(lo + FE_QW, KW)]
        taps[u, :, : seg.shape[1]] = seg
    return {"taps": torch.from_numpy(taps), "base": torch.from_numpy(base), "window_len": int(base.max()) + FE_QW,
            "P": P, "KW": KW}


def pack_mel_csr(fb: torch.Tensor) -> Dict[str, torch.Tensor]:
    """fb [n_freqs, n_mels] -> CSR over mel bands (values, bins, row starts)."""
    f = fb.detach().to("cpu", torch.float32).numpy()
    vals, bins, start = [], [], [0]
    for m in range(f.shape[1]):
        nz = np.nonzero(f[:, m])[0]
        vals.extend(f[nz, m].tolist())
        bins.extend(nz.tolist())
        start.append(len(vals))
    return {"val": torch.tensor(vals, dtype=torch.float32), "bin": torch.tensor(bins, dtype=torch.int32),
            "start": torch.tensor(start, dtype=torch.int32)}


def fft_twiddles(n: int = 1000) -> torch.Tensor:
    k = np.arange(n, dtype=np.float64)
    w = np.exp(-2j * np.pi * k / n)
    return torch.from_numpy(np.stack([w.real, w.imag], axis=1).astype(np.float32))
