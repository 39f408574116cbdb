"""Host-side mirror of the reference's Go API for the Monte-Carlo path (package main of
metropolisMethod/: datatypes.go, metropolis.go, validateResults.go, parsing.go).

Same function names, argument order/meaning and error behaviour as the Go functions, so the parity
tests read like metropolis_test.go.  The heavy functions call the CUDA library through the C ABI
(capi.py); the scalar helpers stay on the host exactly as SURVEY.md 8(b) prescribes ("a kernel
launch for them would be absurd").  The Go twin of this file is go/metropolis_cuda.go.

Aliasing: Go's Molecule copies share their backing array, and JitterLigand /
RotateLigand / RandomizeLigandPose / ShiftLigandCloserByThreshold mutate the caller's atoms in
place.  Molecule here wraps one numpy (n,4) xyzq array; the same functions mutate it in place and
return the same object.
"""
from __future__ import annotations

import math
import os
import random as _random
import sys
from dataclasses import dataclass
from pathlib import Path

import numpy as np

from . import capi

# datatypes.go:6-12
K = 9e9
THRESHOLD = 5
MINDISTANCE = 0.5
MAXANGLE = 0.79
TEMPERATURE = 310.15


@dataclass
class Position3d:                      # datatypes.go:23-25
    X: float = 0.0
    Y: float = 0.0
    Z: float = 0.0

    def Magnitude(self) -> float:      # :43
        return math.sqrt(self.X * self.X + self.Y * self.Y + self.Z * self.Z)

    def Normalize(self) -> None:       # :33 (pointer receiver; no-op at zero magnitude)
        mag = self.Magnitude()
        if mag > 0:
            self.X /= mag
            self.Y /= mag
            self.Z /= mag

    def Dot(self, other: "Position3d") -> float:        # :48
        return self.X * other.X + self.Y * other.Y + self.Z * other.Z

    def Add(self, other: "Position3d") -> "Position3d":  # :53
        return Position3d(self.X + other.X, self.Y + other.Y, self.Z + other.Z)

    def Scale(self, scalar: float) -> "Position3d":      # :58
        return Position3d(self.X * scalar, self.Y * scalar, self.Z * scalar)


@dataclass
class Atom:                            # datatypes.go:18-21
    Position: Position3d
    Charge: float = 0.0


class Molecule:                        # datatypes.go:14-16
    """`atoms []Atom` held as one float64 (n,4) array (X,Y,Z,Charge) -- Go's memory layout."""

    def __init__(self, atoms=None, xyzq=None):
        if xyzq is not None:
            self.xyzq = np.ascontiguousarray(xyzq, dtype=np.float64).reshape(-1, 4)
        elif atoms:
            self.xyzq = np.array([[a.Position.X, a.Position.Y, a.Position.Z, a.Charge] for a in atoms], dtype=np.float64)
        else:
            self.xyzq = np.zeros((0, 4))

    @property
    def atoms(self):
        return [Atom(Position3d(*row[:3]), float(row[3])) for row in self.xyzq]

    def __len__(self):
        return self.xyzq.shape[0]


@dataclass
class MultipleLigandSimulationOutput:  # datatypes.go:27-30
    Ligand: list
    Energy: list


# ---------------------------------------------------------------- run-time configuration
class _Config:
    precision = capi.PRECISION_F32     # pair-sum precision inside the chains
    move_mode = capi.MOVE_REFERENCE
    seed = int.from_bytes(os.urandom(8), "little") >> 1   # Go >= 1.20 seeds math/rand randomly
    calls = 0


config = _Config()
_host_rng = _random.Random()


def Seed(seed: int) -> None:
    """Pin the Philox seed of the device chains and the host helper stream (Go: rand.Seed)."""
    config.seed = int(seed)
    config.calls = 0
    _host_rng.seed(seed)


def _rand_float64() -> float:
    return _host_rng.random()


def _next_seed() -> int:
    # successive top-level calls must not replay the same streams (Go's global source advances)
    s = (config.seed + 0x9E3779B97F4A7C15 * config.calls) & 0xFFFFFFFFFFFFFFFF
    config.calls += 1
    return s


def _params(**kw):
    return capi.default_params(precision=config.precision, move_mode=config.move_mode, **kw)


def _ensure_init():
    if capi.num_devices() == 0:
        capi.init()


def _warn_if_truncated(where):
    """A chain that needed more than max_retries collision retries in one step stopped there; the reference would still be
    retrying (metropolis.go:155, :173): say so instead of returning a truncated chain silently."""
    bits, n = capi.last_chain_status()
    if bits & capi.STATUS_RETRY_LIMIT:
        print(f"ddd_mc warning: {n} chain(s) of {where} stopped at the collision-retry bound; the reference would still be "
              "retrying there (metropolis.go:173)", file=sys.stderr)


# ---------------------------------------------------------------- metropolis.go, device-backed
def SimulateMultipleLigands(protein, ligands, iterations, rotate, temperature, numProcs):
    """metropolis.go:11-22 (serial variant: ShiftLigandCloserByThreshold pre-pass, then per ligand)."""
    if numProcs <= 0:
        raise ZeroDivisionError("integer divide by zero")        # :97
    _ensure_init()
    offsets, xyzq = capi.pack_ligands([l.xyzq for l in ligands])
    seed = _next_seed()
    with capi.Receptor(protein.xyzq, cached=True) as rec:
        scr = capi.Screen(rec, offsets, xyzq, numProcs)
        try:                                                     # one upload: shift, chains, replica pick and gather stay resident
            scr.shift_closer(THRESHOLD)                          # :14-16
            shifted = scr.download_start()
            for i, ligand in enumerate(ligands):                 # the reference shifts the caller's atoms in place (alias)
                ligand.xyzq[:] = shifted[offsets[i]:offsets[i + 1]]
            scr.run(iterations // numProcs, rotate, temperature, _params(), seed=seed)      # :17-18
            out, energy, _, _ = scr.pick(temperature, seed=seed)                            # :102-109, :19
        finally:
            scr.close()
    return [Molecule(xyzq=x) for x in capi.unpack_ligands(offsets, out)], [float(e) for e in energy]


def SimulateMultipleLigandsParallel(protein, ligands, iterations, rotate, temperature, numProcs):
    """metropolis.go:27-51.  One batched device call replaces the numProcs x numProcs goroutine fan-out;
    outputs come back in input order, as the reference's block gather does (:45-49)."""
    if numProcs <= 0:
        raise ZeroDivisionError("integer divide by zero")        # Go: len(ligands)/numProcs panics (:34)
    _ensure_init()
    offsets, xyzq = capi.pack_ligands([l.xyzq for l in ligands])
    with capi.Receptor(protein.xyzq, cached=True) as rec:
        out, energy, _ = capi.minimize_batch(rec, offsets, xyzq, iterations, rotate, temperature, numProcs,
                                             _params(), seed=_next_seed())
    _warn_if_truncated("SimulateMultipleLigandsParallel")
    return [Molecule(xyzq=x) for x in capi.unpack_ligands(offsets, out)], [float(e) for e in energy]


def SimulateLigandMinimizationOneProc(protein, ligands, iterations, rotate, temperature, numProcs, ligandChannel):
    """metropolis.go:56-67; ligandChannel is any object with put() (queue.Queue) or append()."""
    lig, en = SimulateMultipleLigandsParallel(protein, ligands, iterations, rotate, temperature, numProcs)
    out = MultipleLigandSimulationOutput(Ligand=lig, Energy=en)
    (ligandChannel.put if hasattr(ligandChannel, "put") else ligandChannel.append)(out)


def SimulateEnergyMinimizationParallel(protein, ligand, iterations, rotate, temperature, numProcs):
    """metropolis.go:94-111: numProcs replica chains of iterations/numProcs steps + sequential pick."""
    if numProcs <= 0:
        raise ZeroDivisionError("integer divide by zero")        # :97
    lig, _ = SimulateMultipleLigandsParallel(protein, [ligand], iterations, rotate, temperature, numProcs)
    return lig[0]


def _one_chain(protein, ligand, iterations, rotate, temperature):
    _ensure_init()
    offsets, xyzq = capi.pack_ligands([ligand.xyzq])
    with capi.Receptor(protein.xyzq, cached=True) as rec:
        poses, stats, _ = capi.run_chains(rec, offsets, xyzq, 1, iterations, rotate, temperature, _params(),
                                          seed=_next_seed())
    _warn_if_truncated("SimulateEnergyMinimizationOneProc")
    return Molecule(xyzq=poses)


def SimulateEnergyMinimization(protein, ligand, iterations, rotate, temperature):
    """metropolis.go:72-89.  Dead code in the reference (zero callers) whose alias bug never reverts a
    rejected move; kept for API parity and implemented as one correct chain."""
    return _one_chain(protein, ligand, iterations, rotate, temperature)


def SimulateEnergyMinimizationOneProc(protein, ligand, iterations, rotate, temperature, c):
    """metropolis.go:116-136; c is any object with put() or append()."""
    out = _one_chain(protein, ligand, iterations, rotate, temperature)
    (c.put if hasattr(c, "put") else c.append)(out)


def CalculateEnergy(protein, ligand) -> float:
    """metropolis.go:247-262, evaluated on device in fp64 with the reference's per-term arithmetic."""
    _ensure_init()
    offsets, xyzq = capi.pack_ligands([ligand.xyzq])
    with capi.Receptor(protein.xyzq, cached=True) as rec:
        e = capi.energy_batch(rec, offsets, xyzq, capi.default_params(precision=capi.PRECISION_F64))
    return float(e[0])


def FindClosestAtomDistance(ligand, protein):
    """metropolis.go:290-304 -> (distance, ligand atom Position3d, protein atom Position3d)."""
    _ensure_init()
    offsets, xyzq = capi.pack_ligands([ligand.xyzq])
    with capi.Receptor(protein.xyzq, cached=True) as rec:
        d, li, pj = capi.closest_atom_batch(rec, offsets, xyzq)
    lpos = Position3d(*ligand.xyzq[li[0], :3]) if li[0] >= 0 else Position3d()
    ppos = Position3d(*protein.xyzq[pj[0], :3]) if pj[0] >= 0 else Position3d()
    return float(d[0]), lpos, ppos


def ShiftLigandCloserByThreshold(ligand, protein, threshold):
    """metropolis.go:267-285; mutates the ligand's atoms in place and returns the same Molecule."""
    _ensure_init()
    offsets, xyzq = capi.pack_ligands([ligand.xyzq])
    with capi.Receptor(protein.xyzq, cached=True) as rec:
        ligand.xyzq[:] = capi.shift_closer_batch(rec, offsets, xyzq, threshold)
    return ligand


# ---------------------------------------------------------------- metropolis.go, host scalar helpers
def AcceptMove(currentEnergy, newEnergy, temperature, rand=None) -> bool:
    """metropolis.go:141-149: strict '<' short-circuit without a draw, else one draw."""
    if newEnergy < currentEnergy:
        return True
    deltaE = newEnergy - currentEnergy
    probability = math.exp(-deltaE / temperature)
    return (rand or _rand_float64)() < probability


def RotateAtom(pos: Position3d, axis: Position3d, theta: float) -> Position3d:
    """metropolis.go:213-224."""
    cosTheta = math.cos(theta)
    sinTheta = math.sin(theta)
    dot = pos.Dot(axis)
    x = pos.X * cosTheta + sinTheta * (axis.Y * pos.Z - axis.Z * pos.Y) + axis.X * dot * (1 - cosTheta)
    y = pos.Y * cosTheta + sinTheta * (axis.Z * pos.X - axis.X * pos.Z) + axis.Y * dot * (1 - cosTheta)
    z = pos.Z * cosTheta + sinTheta * (axis.X * pos.Y - axis.Y * pos.X) + axis.Z * dot * (1 - cosTheta)
    return Position3d(x, y, z)


def RotateLigand(ligand, maxAngle, rand=None):
    """metropolis.go:189-208; in place (alias), returns the same Molecule."""
    r = rand or _rand_float64
    axis = Position3d(r() * 2 - 1, r() * 2 - 1, r() * 2 - 1)
    axis.Normalize()
    theta = (r() * 2 - 1) * maxAngle
    for row in ligand.xyzq:
        p = RotateAtom(Position3d(row[0], row[1], row[2]), axis, theta)
        row[0], row[1], row[2] = p.X, p.Y, p.Z
    return ligand


def _jitter(ligand, width, r):
    for row in ligand.xyzq:
        row[0] += (r() - 0.5) * width
        row[1] += (r() - 0.5) * width
        row[2] += (r() - 0.5) * width


def JitterLigand(ligand, minDistance, rand=None):
    """metropolis.go:154-167; cumulative retry on the already-moved atoms."""
    r = rand or _rand_float64
    while True:
        _jitter(ligand, 0.1, r)
        if IsCollisionFree(ligand, minDistance):
            return ligand


def JitterAndRotateLigand(ligand, minDistance, maxAngle, rand=None):
    """metropolis.go:172-184."""
    r = rand or _rand_float64
    while True:
        RotateLigand(ligand, maxAngle, r)
        _jitter(ligand, 0.1, r)
        if IsCollisionFree(ligand, minDistance):
            return ligand


def IsCollisionFree(ligand, minDistance) -> bool:
    """metropolis.go:229-242."""
    a = ligand.xyzq
    n = a.shape[0]
    for i in range(n):
        for j in range(i + 1, n):
            dx = a[i, 0] - a[j, 0]
            dy = a[i, 1] - a[j, 1]
            dz = a[i, 2] - a[j, 2]
            if math.sqrt(dx * dx + dy * dy + dz * dz) < minDistance:
                return False
    return True


def Distance(a: Position3d, b: Position3d) -> float:
    """metropolis.go:309-314."""
    dx = a.X - b.X
    dy = a.Y - b.Y
    dz = a.Z - b.Z
    return math.sqrt(dx * dx + dy * dy + dz * dz)


def CopyLigand(ligand):
    """metropolis.go:319-334: deep copy."""
    return Molecule(xyzq=ligand.xyzq.copy())


# ---------------------------------------------------------------- validateResults.go
def CalculateRMSD(simulated, reference) -> float:
    """validateResults.go:50-65 on device; raises IndexError when reference is shorter (Go panics)."""
    if len(reference) < len(simulated):
        raise IndexError("index out of range")
    _ensure_init()
    n = len(simulated)
    out = capi.rmsd_batch(np.array([0, n], dtype=np.int64), simulated.xyzq, reference.xyzq[:n], 4)
    return float(out[0])


def RandomizeLigandPose(ligand, rand=None):
    """validateResults.go:70-79; in place, returns the same Molecule."""
    r = rand or _rand_float64
    _jitter(ligand, 10.0, r)
    return RotateLigand(ligand, math.pi, r)


def CompareRMSD(protein, ligand, iterations, rotate, temperature, numProcs) -> float:
    """validateResults.go:40-45."""
    if numProcs <= 0:
        raise ZeroDivisionError("integer divide by zero")
    _ensure_init()
    offsets, xyzq = capi.pack_ligands([ligand.xyzq])
    seed = _next_seed()
    with capi.Receptor(protein.xyzq, cached=True) as rec:
        scr = capi.Screen(rec, offsets, xyzq, numProcs)          # the uploaded pose stays resident as the reference (:41)
        try:
            scr.randomize(seed=seed)                             # :42
            scr.run(iterations // numProcs, rotate, temperature, _params(), seed=seed)       # :43
            _, _, _, rmsd = scr.pick(temperature, seed=seed)     # :44 CalculateRMSD(simulated, reference)
        finally:
            scr.close()
    return float(rmsd[0])


def average(arr) -> float:
    """validateResults.go:84-93."""
    if len(arr) == 0:
        return 0
    return sum(arr) / float(len(arr))


# ---------------------------------------------------------------- parsing.go
def ParseMol2(filename) -> Molecule:
    """parsing.go:50-106: ATOM section, >= 9 whitespace fields, x/y/z = fields 2..4, charge = last;
    unparsable numbers become 0 (the reference discards strconv errors)."""
    def num(tok):
        try:
            return float(tok)
        except ValueError:
            return 0.0
    rows = []
    in_atoms = False
    with open(filename, "r") as fh:          # raises OSError like the reference's os.Open + Check
        for line in fh:
            line = line.rstrip("\r\n")
            if line.startswith("@<TRIPOS>ATOM"):
                in_atoms = True
                continue
            if in_atoms and line.startswith("@<TRIPOS>"):
                break
            if not in_atoms:
                continue
            fields = line.split()
            if len(fields) < 9:
                continue
            rows.append([num(fields[2]), num(fields[3]), num(fields[4]), num(fields[-1])])
    return Molecule(xyzq=np.array(rows, dtype=np.float64).reshape(-1, 4))


def ParseMol2Batch(filenames, threads=0):
    """Screen-scale ingest: the files parsed on a thread pool straight into the C ABI's CSR staging buffers ->
    (offsets int64[n+1], xyzq float64[sum, 4]), same order as `filenames`."""
    from concurrent.futures import ThreadPoolExecutor
    if not filenames:
        return np.zeros(1, dtype=np.int64), np.zeros((0, 4))
    with ThreadPoolExecutor(max_workers=threads or (os.cpu_count() or 1)) as pool:
        mols = list(pool.map(lambda f: ParseMol2(f).xyzq, filenames))
    return capi.pack_ligands(mols)


def UpdateMol2Coordinates(originalFile, updatedFile, newMolecule) -> None:
    """plotting.go:126-189: rewrite the ATOM section's x/y/z ("%.4f") with the molecule's coordinates, fields re-joined by single
    spaces (strings.Join), every other line copied; raises when the atom counts differ (nothing is written then: the
    reference returns before Flush)."""
    with open(originalFile, "r", newline="") as fh:
        lines = fh.read().split("\n")
    if lines and lines[-1] == "":
        lines.pop()
    out, atomIndex, inAtomSection = [], 0, False
    n = len(newMolecule)
    for line in lines:
        line = line[:-1] if line.endswith("\r") else line          # bufio.ScanLines drops a trailing \r
        if line.startswith("@<TRIPOS>ATOM"):
            inAtomSection = True
            out.append(line)
            continue
        elif line.startswith("@<TRIPOS>") and inAtomSection:
            inAtomSection = False
        if inAtomSection and atomIndex < n:
            parts = line.split()
            if len(parts) >= 9:
                x, y, z = newMolecule.xyzq[atomIndex, :3]
                parts[2], parts[3], parts[4] = "%.4f" % x, "%.4f" % y, "%.4f" % z
                out.append(" ".join(parts))
                atomIndex += 1
            else:
                out.append(line)
        else:
            out.append(line)
    if atomIndex != n:
        open(updatedFile, "w").close()                           # os.Create ran before the scan: an empty file is left behind
        raise ValueError("mismatch between number of atoms in molecule and file")
    with open(updatedFile, "w", newline="") as fh:
        fh.write("".join(l + "\n" for l in out))


def SaveMinimumEnergyLigand(energyList, ligandFiles, fileName, minLigands) -> int:
    """plotting.go:109-121 (minEnergy starts at 0.0: only a negative energy can displace index 0); returns minIndex."""
    minIndex, minEnergy = 0, 0.0
    for i, energy in enumerate(energyList):
        if energy < minEnergy:
            minIndex, minEnergy = i, energy
    UpdateMol2Coordinates(ligandFiles[minIndex], fileName, minLigands[minIndex])
    print(f"Wrote min ligand file to {fileName}", end="")
    return minIndex


def ParseMol2Types(filename) -> list:
    """EXTENSION (LJ): the SYBYL atom-type column (field 5) that parsing.go:84-96 reads past.  Same line filter as ParseMol2,
    so ParseMol2Types(f)[i] is the type of ParseMol2(f).atoms[i]."""
    types = []
    in_atoms = False
    with open(filename, "r") as fh:
        for line in fh:
            line = line.rstrip("\r\n")
            if line.startswith("@<TRIPOS>ATOM"):
                in_atoms = True
                continue
            if in_atoms and line.startswith("@<TRIPOS>"):
                break
            if not in_atoms:
                continue
            fields = line.split()
            if len(fields) < 9:
                continue
            types.append(fields[5])
    return types


# EXTENSION (LJ): 12-6 parameters per SYBYL element (sigma in A, epsilon in kcal/mol; OPLS-AA-like united values).  The
# reference has no such table -- it has no Lennard-Jones term; this only makes ddd_receptor_set_lj /
# ddd_screen_set_lj usable from mol2 files.  Keyed by the element part of the type ("C.ar" -> "C").
LJ_BY_ELEMENT = {
    "C": (3.50, 0.066), "N": (3.25, 0.170), "O": (3.00, 0.170), "S": (3.55, 0.250), "P": (3.74, 0.200), "H": (2.50, 0.030),
    "F": (2.95, 0.053), "Cl": (3.40, 0.300), "Br": (3.47, 0.470), "I": (3.80, 0.600), "Mg": (1.64, 0.875),
    "Zn": (1.96, 0.250), "Ca": (2.41, 0.450), "Na": (3.33, 0.003), "K": (4.93, 0.0003), "Fe": (2.59, 0.013),
}
LJ_DEFAULT = (3.40, 0.100)


def lj_parameters(types):
    """(sigma[n], epsilon[n]) float64 for a list of SYBYL atom types; unknown elements get LJ_DEFAULT."""
    out = np.array([LJ_BY_ELEMENT.get(t.split(".")[0].capitalize() if len(t.split(".")[0]) > 1 else t.split(".")[0].upper(), LJ_DEFAULT)
                    for t in types], dtype=np.float64).reshape(-1, 2)
    return np.ascontiguousarray(out[:, 0]), np.ascontiguousarray(out[:, 1])


def findFilesWithSubstring(rootDir, searchString):
    """parsing.go:142-160: filepath.Walk order (lexical), files whose NAME contains the substring."""
    out = []
    for dirpath, dirnames, filenames in os.walk(rootDir):
        dirnames.sort()
        for name in sorted(filenames):
            if searchString in name:
                out.append(str(Path(dirpath) / name))
    return sorted(out)


def ExtractFileLabel(path) -> str:
    """plotting.go:95-104 (label = base name up to the first '_', ".mol2" suffix trimmed)."""
    label = Path(path).name.split("_")[0]
    return label[:-5] if label.endswith(".mol2") else label
