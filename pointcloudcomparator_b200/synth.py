"""Seeded synthetic clouds for the five BASELINE.json configs.

All generators return float32 arrays in metres; `stride4=True` returns [n, 4] rows (x, y, z, 1.0)
which is the first 16 bytes of pcl::PointXYZ / PointXYZRGB (the reference's point layout,
include/comparator.h:15-46 pulls pcl/point_types.h).  Points are shuffled so input order carries no locality.
"""
from __future__ import annotations

import numpy as np


def _pad4(p, stride4):
    p = np.ascontiguousarray(p, np.float32)
    if not stride4:
        return p
    out = np.ones((p.shape[0], 4), np.float32)
    out[:, :3] = p
    return out


def _box_surface(rng, n, lo, hi):
    lo, hi = np.asarray(lo, np.float64), np.asarray(hi, np.float64)
    e = hi - lo
    areas = np.array([e[1] * e[2], e[1] * e[2], e[0] * e[2], e[0] * e[2], e[0] * e[1], e[0] * e[1]])
    face = rng.choice(6, size=n, p=areas / areas.sum())
    p = lo + rng.random((n, 3)) * e
    axis = face // 2
    side = face % 2
    p[np.arange(n), axis] = np.where(side == 0, lo[axis], hi[axis])
    return p


def _sphere_surface(rng, n, c, r):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return np.asarray(c) + r * v


def room(n: int, seed: int, size=(6.0, 4.0, 2.7), stride4: bool = False, shuffle: bool = True):
    """S1 / S2 / S4: box-shell room + 3 box 'furniture' + 1 sphere, area-uniform surface sampling."""
    rng = np.random.default_rng(seed)
    sx, sy, sz = size
    parts = [
        ("box", (0, 0, 0), (sx, sy, sz)),
        ("box", (0.10 * sx, 0.10 * sy, 0.0), (0.35 * sx, 0.30 * sy, 0.30 * sz)),
        ("box", (0.55 * sx, 0.15 * sy, 0.0), (0.80 * sx, 0.45 * sy, 0.28 * sz)),
        ("box", (0.20 * sx, 0.60 * sy, 0.0), (0.45 * sx, 0.90 * sy, 0.70 * sz)),
        ("sph", (0.70 * sx, 0.70 * sy, 0.25 * sz), 0.22 * sz),
    ]
    areas = []
    for kind, a, b in parts:
        if kind == "box":
            e = np.asarray(b, np.float64) - np.asarray(a, np.float64)
            areas.append(2 * (e[0] * e[1] + e[1] * e[2] + e[0] * e[2]))
        else:
            areas.append(4 * np.pi * b * b)
    areas = np.asarray(areas)
    counts = np.floor(n * areas / areas.sum()).astype(np.int64)
    counts[0] += n - counts.sum()
    chunks = []
    for (kind, a, b), c in zip(parts, counts):
        chunks.append(_box_surface(rng, int(c), a, b) if kind == "box" else _sphere_surface(rng, int(c), a, b))
    p = np.concatenate(chunks).astype(np.float32)
    if shuffle:
        p = p[np.random.default_rng(seed + 7).permutation(n)]
    return _pad4(p, stride4)


def noisy_copy(p, seed: int, sigma: float, every: int = 1, stride4: bool = False):
    """B = A[::every] + N(0, sigma) per coordinate (S1 copy: 2 mm; S2 copy: every 2nd, 5 mm)."""
    rng = np.random.default_rng(seed)
    q = np.asarray(p, np.float32)[::every, :3]
    q = (q + rng.normal(0.0, sigma, q.shape)).astype(np.float32)
    return _pad4(q, stride4)


def rgb_for(p, seed: int):
    """per-surface colour +- 4 uint8 jitter, keyed on the dominant axis-aligned plane; returns uint8 [n,3]."""
    rng = np.random.default_rng(seed)
    p = np.asarray(p)[:, :3]
    key = (np.floor(p[:, 0] * 0.5).astype(np.int64) * 7 + np.floor(p[:, 1] * 0.5).astype(np.int64) * 13 + np.floor(p[:, 2]).astype(np.int64) * 29) % 6
    base = np.array([[200, 60, 60], [60, 200, 60], [60, 60, 200], [200, 200, 60], [60, 200, 200], [200, 60, 200]], np.int64)[key]
    return np.clip(base + rng.integers(-4, 5, base.shape), 0, 255).astype(np.uint8)


def scene(n: int, seed: int, extent: float = 20.0, n_objects: int = 200, gap: float = 0.10, stride4: bool = False):
    """S3: extent x extent field of spheres (r in [0.1, 0.5]) and boxes (edge in [0.2, 1.0]) whose
    bounding spheres are >= gap apart -> the expected Euclidean cluster count equals n_objects
    for any tolerance < gap (and dense enough sampling).  Returns (points, object_id)."""
    rng = np.random.default_rng(seed)
    centres, radii, kinds, sizes = [], [], [], []
    tries = 0
    while len(centres) < n_objects and tries < 200000:
        tries += 1
        kind = rng.integers(0, 2)
        if kind == 0:
            r = rng.uniform(0.1, 0.5)
            br, sz = r, (r,)
        else:
            e = rng.uniform(0.2, 1.0, 3)
            br, sz = 0.5 * float(np.linalg.norm(e)), tuple(e)
        c = np.array([rng.uniform(br, extent - br), rng.uniform(br, extent - br), br])
        ok = True
        for c2, r2 in zip(centres, radii):
            if np.linalg.norm(c - c2) < br + r2 + gap:
                ok = False
                break
        if ok:
            centres.append(c), radii.append(br), kinds.append(kind), sizes.append(sz)
    m = len(centres)
    areas = np.array([4 * np.pi * s[0] ** 2 if k == 0 else 2 * (s[0] * s[1] + s[1] * s[2] + s[0] * s[2]) for k, s in zip(kinds, sizes)])
    counts = np.floor(n * areas / areas.sum()).astype(np.int64)
    counts[0] += n - counts.sum()
    chunks, ids = [], []
    for i in range(m):
        c = int(counts[i])
        if kinds[i] == 0:
            chunks.append(_sphere_surface(rng, c, centres[i], sizes[i][0]))
        else:
            e = np.asarray(sizes[i])
            chunks.append(_box_surface(rng, c, centres[i] - e / 2, centres[i] + e / 2))
        ids.append(np.full(c, i, np.int32))
    p = np.concatenate(chunks).astype(np.float32)
    ids = np.concatenate(ids)
    perm = np.random.default_rng(seed + 7).permutation(p.shape[0])
    return _pad4(p[perm], stride4), ids[perm]


def uniform(n: int, seed: int, extent: float = 10.0, stride4: bool = False):
    """S5a: uniform in [0, extent)^3."""
    rng = np.random.default_rng(seed)
    return _pad4((rng.random((n, 3), dtype=np.float32) * np.float32(extent)), stride4)


def rigid(rz_deg: float = 2.0, t=(0.02, 0.01, 0.005)):
    a = np.deg2rad(rz_deg)
    T = np.eye(4, dtype=np.float64)
    T[0, 0], T[0, 1], T[1, 0], T[1, 1] = np.cos(a), -np.sin(a), np.sin(a), np.cos(a)
    T[:3, 3] = t
    return T


def icp_pair(n: int, seed: int = 4001, size=(10.0, 10.0, 3.0), sigma: float = 0.001, stride4: bool = False):
    """S4: target room; source = R_z(2 deg) T(0.02, 0.01, 0.005) target + N(0, 1 mm).  Returns (source, target, T)."""
    tgt = room(n, seed, size)
    T = rigid()
    rng = np.random.default_rng(seed + 1)
    src = (tgt.astype(np.float64) @ T[:3, :3].T + T[:3, 3] + rng.normal(0, sigma, tgt.shape)).astype(np.float32)
    src = src[np.random.default_rng(seed + 8).permutation(n)]
    return _pad4(src, stride4), _pad4(tgt, stride4), T


def sweep_queries(ref, nq: int, seed: int = 5002, sigma: float = 0.01, stride4: bool = False):
    """S5 queries: random reference points + N(0, 1 cm)."""
    rng = np.random.default_rng(seed)
    ref = np.asarray(ref)[:, :3]
    pick = rng.integers(0, ref.shape[0], nq)
    q = ref[pick] + rng.normal(0, sigma, (nq, 3)).astype(np.float32)
    return _pad4(q.astype(np.float32), stride4)
