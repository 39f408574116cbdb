"""Synthetic receptor / ligand sets of SURVEY.md 8(d) (there is no network for real datasets).

Receptor(nP, seed): points in a ball of radius R = (3 nP / (4 pi 0.035))^(1/3) A by rejection sampling with
minimum separation 1.2 A; charges clip(N(0, 0.2), -0.55, 0.55); everything rounded to 4 decimals (mol2
precision).  Ligand(nL, seed_i): self-avoiding random walk, bond 1.5 A, minimum intra-ligand distance
1.2 A (> MINDISTANCE), charges N(0, 0.18) shifted to net 0 (even index) or -1 (odd index), placed outside
the receptor along a random direction and pulled in the way ShiftLigandCloserByThreshold would
(metropolis.go:267-285) so that the closest receptor atom is ~5 A away.
Deterministic for a given (numpy version, seed).
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree

RECEPTOR_SEED = 20260001
LIGAND_SEED0 = 20260100
DENSITY = 0.035


def receptor_radius(n_atoms: int) -> float:
    return (3.0 * n_atoms / (4.0 * np.pi * DENSITY)) ** (1.0 / 3.0)


def make_receptor(n_atoms: int, seed: int = RECEPTOR_SEED, min_sep: float = 1.2) -> np.ndarray:
    """(n_atoms, 4) float64 xyzq."""
    rng = np.random.default_rng(seed)
    R = receptor_radius(n_atoms)
    pts = np.zeros((0, 3))
    while len(pts) < n_atoms:
        m = max(256, int(1.6 * (n_atoms - len(pts))))
        cand = rng.uniform(-R, R, size=(m, 3))
        cand = cand[(cand ** 2).sum(1) <= R * R]
        if len(pts):
            d, _ = cKDTree(pts).query(cand, k=1)
            cand = cand[d >= min_sep]
        if len(cand) > 1:                       # greedy: drop the later point of every close pair
            drop = np.zeros(len(cand), dtype=bool)
            for i, j in sorted(cKDTree(cand).query_pairs(min_sep)):
                if not drop[i]:
                    drop[j] = True
            cand = cand[~drop]
        pts = np.concatenate([pts, cand])[:n_atoms]
    q = np.clip(rng.normal(0.0, 0.2, size=n_atoms), -0.55, 0.55)
    # a receptor-like offset so coordinates are not centred on the origin (shipped 223l spans 15..51 A)
    pts = pts + np.array([30.0, 15.0, 10.0])
    return np.round(np.concatenate([pts, q[:, None]], axis=1), 4)


def _walk(rng, n: int, bond: float = 1.5, min_sep: float = 1.2) -> np.ndarray:
    pts = [np.zeros(3)]
    while len(pts) < n:
        for _ in range(200):
            v = rng.normal(size=3)
            v *= bond / np.linalg.norm(v)
            p = pts[-1] + v
            if all(np.linalg.norm(p - o) >= min_sep for o in pts):
                pts.append(p)
                break
        else:                                   # dead end: restart the walk
            pts = [np.zeros(3)]
    return np.array(pts)


def make_ligand(n_atoms: int, index: int, receptor: np.ndarray, tree: cKDTree | None = None,
                seed0: int = LIGAND_SEED0) -> np.ndarray:
    """(n_atoms, 4) float64 xyzq for ligand `index` (seed = seed0 + index)."""
    rng = np.random.default_rng(seed0 + index)
    xyz = _walk(rng, n_atoms)
    q = rng.normal(0.0, 0.18, size=n_atoms)
    q += ((0.0 if index % 2 == 0 else -1.0) - q.sum()) / n_atoms
    centre = receptor[:, :3].mean(0)
    R = np.linalg.norm(receptor[:, :3] - centre, axis=1).max()
    u = rng.normal(size=3)
    u /= np.linalg.norm(u)
    xyz = xyz - xyz.mean(0) + centre + u * (R + 15.0)
    tree = tree or cKDTree(receptor[:, :3])
    for _ in range(3):                          # ShiftLigandCloserByThreshold-style pull-in
        d, j = tree.query(xyz, k=1)
        i = int(np.argmin(d))
        if d[i] <= 5.0:
            break
        v = receptor[j[i], :3] - xyz[i]
        xyz = xyz + v / np.linalg.norm(v) * (d[i] - 5.0)
    return np.round(np.concatenate([xyz, q[:, None]], axis=1), 4)


_POOL_STATE = {}


def _pool_init(receptor, seed0):
    _POOL_STATE["rec"] = receptor
    _POOL_STATE["tree"] = cKDTree(receptor[:, :3])
    _POOL_STATE["seed0"] = seed0


def _pool_make(job):
    n, index = job
    return make_ligand(n, index, _POOL_STATE["rec"], _POOL_STATE["tree"], _POOL_STATE["seed0"])


def make_ligands(sizes, receptor: np.ndarray, first_index: int = 0, seed0: int = LIGAND_SEED0, workers: int = 1):
    """CSR batch: (offsets int64[n+1], xyzq float64[sum, 4]) for ligands first_index.. with the given sizes.
    Every ligand has its own seed, so `workers` > 1 (a process pool; call it before CUDA is initialised) gives the
    same bits as the serial loop."""
    sizes = [int(n) for n in sizes]
    if workers > 1 and len(sizes) >= 256:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(workers, initializer=_pool_init, initargs=(receptor, seed0)) as pool:
            ligs = pool.map(_pool_make, [(n, first_index + i) for i, n in enumerate(sizes)], chunksize=64)
    else:
        tree = cKDTree(receptor[:, :3])
        ligs = [make_ligand(n, first_index + i, receptor, tree, seed0) for i, n in enumerate(sizes)]
    offsets = np.zeros(len(ligs) + 1, dtype=np.int64)
    offsets[1:] = np.cumsum([len(l) for l in ligs])
    xyzq = np.concatenate(ligs, axis=0) if ligs else np.zeros((0, 4))
    return offsets, np.ascontiguousarray(xyzq)


def sweep_sizes(n_ligands: int, lo: int = 20, hi: int = 80, seed: int = 20260002) -> np.ndarray:
    """Config 5: nL ~ U{lo..hi}."""
    return np.random.default_rng(seed).integers(lo, hi + 1, size=n_ligands)
