"""Shared fixtures/data generators for the test-suite (no GPU needed to import)."""
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
EMPTY = np.uint64(0xFFFFFFFFFFFFFFFF)


def nine_point():
    """9-point dataset + golden bottom layer of the reference's unit tests."""
    with open(os.path.join(GOLDEN, "nine_point.json")) as f:
        g = json.load(f)
    s = np.float32(0.70710678118654752440)  # std::f32::consts::FRAC_1_SQRT_2

    def conv(row):
        return [s if x == "s" else np.float32(x) for x in row]

    g["rows"] = np.array([conv(r) for r in g["data"]], dtype=np.float32)
    g["query"] = np.array(conv(g["nearness_query"]), dtype=np.float32)
    g["neighbors"] = np.array(g["bottom_neighbors"], dtype=np.uint64).reshape(9, 6)
    return g


def nine_point_layers(entry=0):
    """Two-layer stack as Hnsw::generate shapes it for N=9, order=6: a one-node top layer
    (M=3, all empty) above the golden bottom layer."""
    g = nine_point()
    top_nodes = np.array([entry], dtype=np.uint64)
    top_neigh = np.full((1, 3), EMPTY, dtype=np.uint64)
    bottom_nodes = np.arange(9, dtype=np.uint64)
    return g, [(top_nodes, top_neigh, 3), (bottom_nodes, g["neighbors"], 6)]


def random_normed(n, dim, seed):
    """bigvec.rs:59-65 shape: Uniform(-1,1)^d then L2-normalised (our own generator)."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1.0, 1.0, size=(n, dim)).astype(np.float32)
    x /= np.sqrt((x * x).sum(axis=1, dtype=np.float32))[:, None]
    return np.ascontiguousarray(x, dtype=np.float32)


def clustered(n, dim, seed, n_clusters=64, spread=0.15, normalise=False, integer=False,
              centers_seed=977):
    """Gaussian-mixture data (SIFT/Deep-shaped configs of SURVEY section 8d).  The mixture
    (centres) is keyed by centers_seed so that rows and queries drawn with different `seed`
    come from the same distribution."""
    centers = np.random.default_rng(centers_seed).normal(size=(n_clusters, dim)).astype(np.float32)
    rng = np.random.default_rng(seed)
    which = rng.integers(0, n_clusters, size=n)
    x = centers[which] + spread * rng.normal(size=(n, dim)).astype(np.float32)
    if integer:
        x = np.clip(np.rint(40.0 * x + 60.0), 0, 218)
    x = x.astype(np.float32)
    if normalise:
        x /= np.sqrt((x * x).sum(axis=1, dtype=np.float32))[:, None]
    return np.ascontiguousarray(x, dtype=np.float32)


def exact_knn(rows, queries, k, metric="l2"):
    """float64 brute force for recall checks (small sizes only)."""
    r = rows.astype(np.float64)
    q = queries.astype(np.float64)
    if metric == "l2":
        d = (q * q).sum(1)[:, None] - 2.0 * q @ r.T + (r * r).sum(1)[None, :]
    else:
        d = -(q @ r.T)
    idx = np.argpartition(d, k - 1, axis=1)[:, :k]
    dd = np.take_along_axis(d, idx, axis=1)
    order = np.argsort(dd, axis=1, kind="stable")
    return np.take_along_axis(idx, order, axis=1)
