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


def pack_resample_taps(kernel: torch.Tensor, orig_step: int = 0, role_threads: int = 256, hops_per_group: int = 0) -> Dict[str, object]:
    """resampler.kernel [P,1,KW] -> per phase-quad taps over the quad's common window (kernel ABI), plus the
    thread -> (quad, hop slice) map of the resample role that keeps its shared-memory reads bank-conflict free.

    A resample thread owns one phase quad and the hops h = slice (mod n_slices) of its frame group; at tap j it reads the staged
    signal at word  tap_base[quad] + h * O + j.  The 32 lanes of a warp therefore hit distinct banks iff their keys
    (tap_base[quad] + slice * O) mod 32 are distinct.  The map spreads the (quad, slice) work items over the role's warps so that
    this holds (each key value occurs at most once per warp); quads whose non-zero taps span fewer than FE_QW columns may start
    their window one column early to thin out an over-full key.  Inside a warp, lanes are ordered so that every aligned group of 8
    lanes has distinct (quad mod 8): their 16-byte frame stores then fall into distinct bank groups."""
    k = kernel.detach().to("cpu", torch.float32)[:, 0, :].numpy()
    P, KW = k.shape
    if P % 4:
        raise NotImplementedError(f"a number of resample phases that is not a multiple of 4 ({P}) is not supported by the kernel")
    nq = P // 4
    O = int(orig_step)          # input samples per hop (orig_rate / gcd); 0: no thread map (the kernel falls back to its default)
    lo_hi = []
    for u in range(nq):
        nz = np.nonzero((k[4 * u: 4 * u + 4] != 0).any(axis=0))[0]
        lo, hi = (int(nz[0]), int(nz[-1]) + 1) if nz.size else (0, 1)
        if hi - lo > FE_QW:
            raise NotImplementedError(
                f"resample phase quad {u} spans {hi - lo} taps > {FE_QW}: this sample-rate pair is not supported")
        lo_hi.append((lo, hi))
    base = np.array([lo for lo, _ in lo_hi], np.int32)

    # ---- thread map of the resample role
    n_warps = role_threads // 32
    n_slices = max(1, role_threads // nq)
    if hops_per_group:
        n_slices = min(n_slices, hops_per_group)
    lane_map = np.full((role_threads,), -1, np.int32)
    if O > 0 and nq * n_slices <= role_threads:
        def key_hist(b):
            h = np.zeros(32, np.int64)
            for q in range(nq):
                for s_ in range(n_slices):
                    h[(int(b[q]) + s_ * O) % 32] += 1
            return h
        for _ in range(4 * nq):                       # thin out over-full keys with the quads that have a spare column
            h = key_hist(base)
            if h.max() <= n_warps:
                break
            best = None
            for q in range(nq):
                lo, hi = lo_hi[q]
                if base[q] != lo or hi - lo >= FE_QW or lo == 0:
                    continue
                gain = sum(h[(lo + s_ * O) % 32] > n_warps for s_ in range(n_slices))
                b2 = base.copy()
                b2[q] = lo - 1
                h2 = key_hist(b2)
                score = (int((np.maximum(h2 - n_warps, 0)).sum()), -gain)
                if best is None or score < best[0]:
                    best = (score, q)
            if best is None or best[0][0] >= int(np.maximum(h - n_warps, 0).sum()):
                break
            base[best[1]] -= 1
        by_key: Dict[int, list] = {}
        for q in range(nq):
            for s_ in range(n_slices):
                by_key.setdefault((int(base[q]) + s_ * O) % 32, []).append((q, s_))
        warps = [[] for _ in range(n_warps)]
        spill = []
        for key in sorted(by_key, key=lambda kk: -len(by_key[kk])):
            order = sorted(range(n_warps), key=lambda w: len(warps[w]))     # emptiest warps first
            items = by_key[key]
            for it, w in zip(items, order):
                warps[w].append(it)
            spill.extend(items[n_warps:])              # key occurs more often than there are warps: a 2-way conflict somewhere
        for it in spill:
            w = min(range(n_warps), key=lambda w_: len(warps[w_]))
            warps[w].append(it)
        # items with equal keys may swap warps freely: local search until no (quad mod 8) class has more than 4 members
        # (= 4 octets) in any warp
        def excess(items):
            c = np.bincount([q % 8 for q, _ in items], minlength=8)
            return int(np.maximum(c - 4, 0).sum())
        rng = np.random.RandomState(0)
        key_of = lambda it: (int(base[it[0]]) + it[1] * O) % 32   # noqa: E731
        for _ in range(20000):
            if sum(excess(w_) for w_ in warps) == 0:
                break
            a, b_ = rng.randint(n_warps), rng.randint(n_warps)
            if a == b_ or not warps[a]:
                continue
            ia = rng.randint(len(warps[a]))
            ka = key_of(warps[a][ia])
            jb = [j for j, it in enumerate(warps[b_]) if key_of(it) == ka]
            before = excess(warps[a]) + excess(warps[b_])
            if jb:
                warps[a][ia], warps[b_][jb[0]] = warps[b_][jb[0]], warps[a][ia]
                if excess(warps[a]) + excess(warps[b_]) > before:
                    warps[a][ia], warps[b_][jb[0]] = warps[b_][jb[0]], warps[a][ia]
            elif len(warps[b_]) < 30:                                   # move (the key is absent from warp b)
                it = warps[a].pop(ia)
                warps[b_].append(it)
                if excess(warps[a]) + excess(warps[b_]) > before:
                    warps[b_].pop()
                    warps[a].insert(ia, it)
        for w, items in enumerate(warps):
            assert len(items) <= 32
            octets = [[] for _ in range(4)]            # 8 lanes each; distinct (quad mod 8) inside an octet where possible
            by_r8: Dict[int, list] = {}
            for it in items:
                by_r8.setdefault(it[0] % 8, []).append(it)
            for r8 in sorted(by_r8, key=lambda r: -len(by_r8[r])):
                for i, it in enumerate(by_r8[r8]):
                    free = [o for o in range(4) if len(octets[o]) < 8]
                    clean = [o for o in free if all(x[0] % 8 != r8 for x in octets[o])]
                    o = min(clean or free, key=lambda o_: len(octets[o_]))
                    octets[o].append(it)
            lanes = []
            for o in octets:
                lanes.extend(o + [None] * (8 - len(o)))
            for ln, it in enumerate(lanes[:32]):
                if it is not None:
                    lane_map[32 * w + ln] = it[0] | (it[1] << 16)
        assert (lane_map >= 0).sum() == nq * n_slices
    else:
        lane_map = None
    taps = np.zeros((nq, 4, FE_QW), np.float32)
    for u in range(nq):
        lo = int(base[u])
        seg = k[4 * u: 4 * u + 4, lo:min(lo + FE_QW, KW)]
        taps[u, :, : seg.shape[1]] = seg
    return {"taps": torch.from_numpy(taps), "base": torch.from_numpy(base), "window_len": int(base.max()) + FE_QW,
            "P": P, "KW": KW, "lane_map": None if lane_map is None else torch.from_numpy(lane_map), "n_slices": n_slices}


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
