"""Golden fixture for the anchor clustering, produced by the LIVE sklearn KMeans exactly the way the reference's
compute_anchors.py runs it (compute_anchors.py:9-13,72-86: numpy global RNG seeded with 42, KMeans(9, init='k-means++',
n_init='auto', tol=1e-10, max_iter=500) on the column of segment durations).

    python tests/golden/make_golden_anchors.py        -> tests/golden/anchors.npz"""
import os

import numpy as np
from sklearn.cluster import KMeans

HERE = os.path.dirname(os.path.abspath(__file__))


def synth_durations(n: int, seed: int) -> np.ndarray:
    """Segment durations like the reference's dataset (contiguous segments tiling 60 s clips): a mixture of short, medium and
    long segments, clipped to (0.2, 60] seconds."""
    g = np.random.RandomState(seed)
    parts = [g.gamma(2.0, 2.0, n // 2), g.normal(25.0, 6.0, n // 4), g.normal(50.0, 5.0, n - n // 2 - n // 4)]
    return np.clip(np.concatenate(parts), 0.2, 60.0)


def main():
    out = {}
    for name, n, seed, init in (("a", 4000, 1, "k-means++"), ("b", 257, 2, "k-means++"), ("c", 1500, 3, "random")):
        d = synth_durations(n, seed)
        np.random.seed(42)                                        # compute_anchors.py:9-13
        km = KMeans(9, init=init, n_init="auto", tol=1e-10, max_iter=500).fit(d.reshape(-1, 1))
        a = np.sort(km.cluster_centers_.reshape(-1))
        out[f"{name}.durations"] = d
        out[f"{name}.anchors"] = a
        out[f"{name}.n_iter"] = np.int64(km.n_iter_)
        out[f"{name}.inertia"] = np.float64(km.inertia_)
        # Lloyd alone from a given start (array init, one run): pins the iteration independently of the seeding
        init_c = np.quantile(d, np.linspace(0.05, 0.95, 9)).reshape(-1, 1)
        km2 = KMeans(9, init=init_c, n_init=1, tol=1e-10, max_iter=500).fit(d.reshape(-1, 1))
        out[f"{name}.lloyd_init"] = init_c.reshape(-1)
        out[f"{name}.lloyd_centers"] = km2.cluster_centers_.reshape(-1)
        out[f"{name}.lloyd_n_iter"] = np.int64(km2.n_iter_)
        out[f"{name}.lloyd_labels"] = km2.labels_.astype(np.int32)
        print(name, init, "n_iter", km.n_iter_, "anchors", np.round(a, 3), "lloyd n_iter", km2.n_iter_)
    np.savez_compressed(os.path.join(HERE, "anchors.npz"), **out)


if __name__ == "__main__":
    main()
