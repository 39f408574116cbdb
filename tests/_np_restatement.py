"""A second, independently written restatement of the reference's arithmetic in numpy / pure Python
(float64), used ONLY to cross-check the C oracle where the reference's own tests pin nothing
(energy magnitudes, move sequences, RMSD, replica pick).  Small cases only: it is slow on purpose
(scalar loops in reference order).  Cites metropolisMethod/<file>:<line>."""
import math

import numpy as np

K = 9e9


def philox4x32_10(ctr, key):
    """Random123 Philox4x32-10."""
    M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
    c = [int(x) & 0xFFFFFFFF for x in ctr]
    k = [int(x) & 0xFFFFFFFF for x in key]
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & 0xFFFFFFFF, p1 & 0xFFFFFFFF,
             ((p0 >> 32) ^ c[3] ^ k[1]) & 0xFFFFFFFF, p0 & 0xFFFFFFFF]
        k = [(k[0] + W0) & 0xFFFFFFFF, (k[1] + W1) & 0xFFFFFFFF]
    return c


def u01(seed, stream, index):
    blk = index >> 1
    o = philox4x32_10([blk & 0xFFFFFFFF, blk >> 32, stream & 0xFFFFFFFF, stream >> 32],
                      [seed & 0xFFFFFFFF, seed >> 32])
    w = (o[2] << 32 | o[3]) if index & 1 else (o[0] << 32 | o[1])
    return (w >> 11) * 2.0 ** -53


class Stream:
    def __init__(self, seed=None, stream=None, recorded=None):
        self.seed, self.stream, self.rec, self.pos = seed, stream, recorded, 0

    def __call__(self):
        i = self.pos
        self.pos += 1
        if self.rec is not None:
            return float(self.rec[i]) if i < len(self.rec) else 0.5
        return u01(self.seed, self.stream, i)


def energy(prot, lig):
    """metropolis.go:247-262, scalar loops in the reference's order."""
    e = 0.0
    for p in prot:
        for l in lig:
            dx, dy, dz = p[0] - l[0], p[1] - l[1], p[2] - l[2]
            d = math.sqrt(dx * dx + dy * dy + dz * dz)                 # math.Pow(x, 2) == x*x
            if d < 1e-6:
                d = 1e-6
            e += K * (p[3] * l[3]) / d
    return e


def energy_vec(prot, lig):
    """Same terms, vectorised (different summation order): for larger cases, ~1e-13 relative."""
    d = np.sqrt(((prot[:, None, :3] - lig[None, :, :3]) ** 2).sum(-1))
    d = np.maximum(d, 1e-6)
    t = K * (prot[:, None, 3] * lig[None, :, 3]) / d
    return float(t.sum()), float(np.abs(t).sum())


def collision_free(lig, min_distance):
    n = len(lig)
    for i in range(n):
        for j in range(i + 1, n):
            dx, dy, dz = lig[i][0] - lig[j][0], lig[i][1] - lig[j][1], lig[i][2] - lig[j][2]
            if math.sqrt(dx * dx + dy * dy + dz * dz) < min_distance:  # :235-236
                return False
    return True


def rotate(lig, max_angle, r):
    """RotateLigand :189-208 / RotateAtom :213-224, in place."""
    ax = [r() * 2 - 1, r() * 2 - 1, r() * 2 - 1]
    mag = math.sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2])
    if mag > 0:
        ax = [a / mag for a in ax]
    th = (r() * 2 - 1) * max_angle
    c, s = math.cos(th), math.sin(th)
    for a in lig:
        x, y, z = a[0], a[1], a[2]
        dot = x * ax[0] + y * ax[1] + z * ax[2]
        a[0] = x * c + s * (ax[1] * z - ax[2] * y) + ax[0] * dot * (1 - c)
        a[1] = y * c + s * (ax[2] * x - ax[0] * z) + ax[1] * dot * (1 - c)
        a[2] = z * c + s * (ax[0] * y - ax[1] * x) + ax[2] * dot * (1 - c)


def propose(lig, rotate_flag, r, min_distance=0.5, max_angle=0.79, width=0.1):
    """JitterAndRotateLigand :172-184 / JitterLigand :154-167; returns rejected attempts."""
    rej = 0
    while True:
        if rotate_flag:
            rotate(lig, max_angle, r)
        for a in lig:
            a[0] += (r() - 0.5) * width
            a[1] += (r() - 0.5) * width
            a[2] += (r() - 0.5) * width
        if collision_free(lig, min_distance):
            return rej
        rej += 1


def accept(cur, new, T, r):
    if new < cur:                                                     # :142
        return True
    return r() < math.exp(-(new - cur) / T)                            # :147-148


def chain(prot, lig, iterations, rotate_flag, T, r):
    """SimulateEnergyMinimizationOneProc :116-136 on a private copy."""
    lig = np.array(lig, dtype=np.float64, copy=True)
    cur = energy(prot, lig)
    trace, acc = [], 0
    for _ in range(iterations):
        prev = lig.copy()
        propose(lig, rotate_flag, r)
        new = energy(prot, lig)
        if accept(cur, new, T, r):
            cur = new
            acc += 1
            trace.append(1)
        else:
            lig = prev
            trace.append(0)
    return lig, cur, trace


def minimize_parallel(prot, lig, iterations, rotate_flag, T, num_procs, seed, lig_index=0):
    """SimulateEnergyMinimizationParallel :94-111, replica-index order, oracle stream convention."""
    cur_e = energy(prot, lig)
    width = iterations // num_procs
    ends = [chain(prot, lig, width, rotate_flag, T, Stream(seed, lig_index * num_procs + rp))[0] for rp in range(num_procs)]
    parent = Stream(seed, 0x4000000000000000 + lig_index)
    cur = np.array(lig, copy=True)
    for e in ends:
        new_e = energy(prot, e)
        if accept(cur_e, new_e, T, parent):
            cur, cur_e = e, new_e
    return cur, cur_e


def rmsd(sim, ref):
    s = 0.0
    for a, b in zip(sim, ref[:len(sim)]):                              # validateResults.go:54-63
        dx, dy, dz = a[0] - b[0], a[1] - b[1], a[2] - b[2]
        s += dx * dx + dy * dy + dz * dz
    return math.sqrt(s / len(sim))


def find_closest(lig, prot):
    best, bi, bj = 1.7976931348623157e308, -1, -1                      # :291
    for i, a in enumerate(lig):
        for j, p in enumerate(prot):
            d = math.sqrt((a[0] - p[0]) * (a[0] - p[0]) + (a[1] - p[1]) * (a[1] - p[1]) + (a[2] - p[2]) * (a[2] - p[2]))
            if d < best:
                best, bi, bj = d, i, j
    return best, bi, bj
