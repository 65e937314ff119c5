"""Anchor clustering: the reference's ``compute_anchors.py`` (KMeans over the annotated segment durations -> 9 sorted centres ->
``anchors.sm / md / lg`` of config.yaml) with the Lloyd iterations on the GPU (``yad_kmeans1d_lloyd``, fp64, one launch).

The host side mirrors what sklearn 1.9's ``KMeans.fit`` does around the iterations (compute_anchors.py:72-86 ->
sklearn/cluster/_kmeans.py): centring, the variance-scaled tolerance, ``n_init`` restarts, and the k-means++ / random seeding
drawn from a ``numpy.random.RandomState`` exactly as sklearn draws it (the reference relies on numpy's global RNG seeded with 42,
compute_anchors.py:9-13), so that a given seed picks the same initial centres.  Seeding is a sequential, RNG-driven selection
over a few thousand scalars and stays on the host; the iterations run in the kernel, never on the CPU."""
from __future__ import annotations

import ctypes as C
import json
from typing import Dict, Optional, Sequence, Tuple

import numpy as np
import torch
import yaml

from . import _lib

NUM_CLUSTERS = 9          # compute_anchors.py:10


def _sq_dists(xc: np.ndarray, x: np.ndarray, xsq: np.ndarray) -> np.ndarray:
    d = -2.0 * (xc[:, None] * x[None, :])
    d += (xc * xc)[:, None]
    d += xsq[None, :]
    np.maximum(d, 0, out=d)
    return d


def _seed_kmeans_plusplus(x: np.ndarray, k: int, rng: np.random.RandomState) -> np.ndarray:
    """sklearn _kmeans_plusplus for unit sample weights on centred 1-D data (same RNG calls in the same order)."""
    n = x.shape[0]
    xsq = x * x
    w = np.ones(n)
    trials = 2 + int(np.log(k))
    centers = np.empty(k)
    centers[0] = x[rng.choice(n, p=w / w.sum())]
    closest = _sq_dists(centers[:1], x, xsq)
    pot = closest @ w
    for c in range(1, k):
        cand = np.searchsorted(np.cumsum(w * closest), rng.uniform(size=trials) * pot)
        np.clip(cand, None, closest.size - 1, out=cand)
        d = _sq_dists(x[cand], x, xsq)
        np.minimum(closest, d, out=d)
        cpot = d @ w.reshape(-1, 1)
        best = int(np.argmin(cpot))
        pot, closest = cpot[best], d[best]
        centers[c] = x[cand[best]]
    return centers


def kmeans_lloyd(x: torch.Tensor, centers_init: Sequence[float], max_iter: int, tol_abs: float) -> Dict[str, object]:
    """Lloyd iterations on the GPU.  x: [n] float64 CUDA tensor (already centred), centers_init: k start centres.
    Returns centres (numpy, float64), labels (CUDA int32), inertia, n_iter."""
    if not x.is_cuda or x.dtype != torch.float64:
        raise RuntimeError("kmeans_lloyd needs a float64 CUDA tensor (no CPU fallback)")
    lib = _lib.init(x.device.index if x.device.index is not None else torch.cuda.current_device())
    x = x.contiguous()
    c = torch.as_tensor(np.asarray(centers_init, np.float64), device=x.device).clone()
    labels = torch.empty(x.numel(), dtype=torch.int32, device=x.device)
    n_iter = torch.zeros(1, dtype=torch.int32, device=x.device)
    inertia = torch.zeros(1, dtype=torch.float64, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.yad_kmeans1d_lloyd(x.data_ptr(), x.numel(), c.data_ptr(), c.numel(), int(max_iter), float(tol_abs), labels.data_ptr(),
                                    n_iter.data_ptr(), inertia.data_ptr(), C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream))
    _lib.check(rc, "kmeans1d_lloyd")
    return {"centers": c.cpu().numpy(), "labels": labels, "inertia": float(inertia.item()), "n_iter": int(n_iter.item())}


def compute_anchors(durations: Sequence[float], n_clusters: int = NUM_CLUSTERS, init: str = "k-means++", n_init="auto",
                    max_iter: int = 500, tol: float = 1e-10, random_state: Optional[np.random.RandomState] = None,
                    device: Optional[torch.device] = None) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """compute_anchors.py:72-86: returns (sm, md, lg) = the sorted cluster centres cut into three triples (seconds)."""
    if init not in ("k-means++", "random"):
        raise ValueError("init must be 'k-means++' or 'random'")
    rng = random_state if random_state is not None else np.random.mtrand._rand      # numpy's global RNG, like sklearn's default
    X = np.asarray(durations, np.float64).reshape(-1)
    if X.shape[0] < n_clusters:
        raise ValueError(f"n_samples={X.shape[0]} should be >= n_clusters={n_clusters}.")
    tol_abs = float(np.var(X)) * tol
    mean = X.mean()
    x = X - mean
    dev = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    xd = torch.from_numpy(x).to(dev)
    if n_init == "auto":
        n_init = 1 if init == "k-means++" else 10
    best = None
    for _ in range(int(n_init)):
        if init == "k-means++":
            c0 = _seed_kmeans_plusplus(x, n_clusters, rng)
        else:
            c0 = x[rng.choice(x.shape[0], size=n_clusters, replace=False, p=np.ones(x.shape[0]) / x.shape[0])]
        r = kmeans_lloyd(xd, c0, max_iter, tol_abs)
        if best is None or r["inertia"] < best["inertia"]:
            best = r
    a = np.sort(best["centers"] + mean)
    k3 = n_clusters // 3
    return a[:k3], a[k3:2 * k3], a[2 * k3:]


def durations_from_annotations(path: str, annotator: str = "annotator_a") -> np.ndarray:
    """compute_anchors.py:60-70: segment durations of a JSON annotation file (plain or grouped)."""
    with open(path, "r") as f:
        ann = json.load(f)["annotations"][annotator]
    out = []
    for v in ann.values():
        for seg in v.values():
            if isinstance(seg, dict) and "start" in seg and "end" in seg:
                out.append(seg["end"] - seg["start"])
            else:                                   # grouped annotations: one more level
                for i in seg.values():
                    out.append(i["end"] - i["start"])
    return np.asarray(out, np.float64)


def set_config_anchors(config_path: str, sm: Sequence[float], md: Sequence[float], lg: Sequence[float]) -> None:
    """compute_anchors.py:21-31: write the three anchor triples back into config.yaml."""
    with open(config_path, "r") as f:
        data = yaml.safe_load(f)
    data["anchors"]["sm"] = np.asarray(sm).tolist()
    data["anchors"]["md"] = np.asarray(md).tolist()
    data["anchors"]["lg"] = np.asarray(lg).tolist()
    with open(config_path, "w") as f:
        yaml.safe_dump(data, f)
