"""ctypes wrapper of the CPU oracle (oracle/libddd_oracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
ORACLE_DIR = ROOT / "oracle"

PARENT_BASE = 0x4000000000000000
RANDOMIZE_BASE = 0x2000000000000000


class OrcParams(C.Structure):
    _fields_ = [
        ("k_coulomb", C.c_double), ("threshold", C.c_double), ("min_distance", C.c_double),
        ("max_angle", C.c_double), ("jitter_width", C.c_double), ("clamp_distance", C.c_double),
        ("move_mode", C.c_int), ("rigid_translation", C.c_double), ("rigid_angle", C.c_double),
        ("max_retries", C.c_int64), ("cutoff", C.c_double),
        ("lj_scale", C.c_double), ("lj_cap", C.c_double), ("prot_lj", C.c_void_p), ("lig_lj", C.c_void_p),
    ]


class OrcStream(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("seed", C.c_uint64), ("stream_id", C.c_uint64), ("pos", C.c_uint64),
        ("rec", C.c_void_p), ("rec_len", C.c_uint64), ("exhausted", C.c_int),
    ]


STATS_DTYPE = np.dtype([
    ("steps", "<i8"), ("accepted", "<i8"), ("downhill", "<i8"), ("retries", "<i8"), ("draws", "<i8"),
    ("status", "<i8"), ("final_energy", "<f8"), ("start_energy", "<f8"),
])


def _build():
    res = subprocess.run(["make", "-C", str(ORACLE_DIR)], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + res.stdout + res.stderr)


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _atoms(a):
    return np.ascontiguousarray(a, dtype=np.float64).reshape(-1, 4)


class Oracle:
    def __init__(self, variant: str = ""):
        name = {"pow": "libddd_oracle_pow.so", "gopow": "libddd_oracle_gopow.so"}.get(variant, "libddd_oracle.so")
        path = ORACLE_DIR / name
        src_m = max((ORACLE_DIR / f).stat().st_mtime for f in ("ddd_oracle.c", "ddd_oracle.h"))
        if not path.exists() or path.stat().st_mtime < src_m:
            _build()
        L = self.L = C.CDLL(str(path))
        d, i64, u64, vp = C.c_double, C.c_int64, C.c_uint64, C.c_void_p
        PP, SP = C.POINTER(OrcParams), C.POINTER(OrcStream)
        sig = {
            "orc_params_default": (None, [PP]),
            "orc_stream_philox": (None, [SP, u64, u64]),
            "orc_stream_recorded": (None, [SP, vp, u64]),
            "orc_stream_next": (d, [SP]),
            "orc_philox4x32_10": (None, [vp, vp, vp]),
            "orc_philox_u01": (d, [u64, u64, u64]),
            "orc_magnitude": (d, [vp]), "orc_normalize": (None, [vp]), "orc_dot": (d, [vp, vp]),
            "orc_distance": (d, [vp, vp]),
            "orc_rotate_atom": (None, [vp, vp, d, vp]),
            "orc_is_collision_free": (C.c_int, [vp, i64, d]),
            "orc_calculate_energy": (d, [vp, i64, vp, i64, PP]),
            "orc_calculate_energy_abs": (d, [vp, i64, vp, i64, PP, vp]),
            "orc_chain_one_proc_best": (None, [vp, i64, vp, i64, i64, C.c_int, d, PP, SP, vp, vp, vp]),
            "orc_min_energy_index": (i64, [vp, i64, C.c_int, vp]),
            "orc_accept_move": (C.c_int, [d, d, d, SP]),
            "orc_rotate_ligand": (None, [vp, i64, d, SP]),
            "orc_jitter_ligand": (i64, [vp, i64, PP, SP]),
            "orc_jitter_and_rotate_ligand": (i64, [vp, i64, PP, SP]),
            "orc_find_closest_atom_distance": (d, [vp, i64, vp, i64, vp, vp]),
            "orc_shift_ligand_closer": (None, [vp, i64, vp, i64, d]),
            "orc_rigid_move": (None, [vp, i64, PP, SP]),
            "orc_chain_one_proc": (None, [vp, i64, vp, i64, i64, C.c_int, d, PP, SP, vp, vp, vp]),
            "orc_minimize_parallel": (C.c_int, [vp, i64, vp, i64, i64, C.c_int, d, C.c_int, PP, u64, u64, vp, vp, vp]),
            "orc_simulate_multiple_ligands": (C.c_int, [vp, i64, vp, vp, i64, i64, C.c_int, d, C.c_int, PP, u64,
                                                        C.c_int, C.c_int, vp, vp]),
            "orc_ligand_block": (None, [i64, C.c_int, C.c_int, vp, vp]),
            "orc_calculate_rmsd": (d, [vp, i64, vp, i64]),
            "orc_randomize_ligand_pose": (None, [vp, i64, SP]),
            "orc_compare_rmsd": (d, [vp, i64, vp, i64, i64, C.c_int, d, C.c_int, PP, u64, u64]),
            "orc_parse_mol2": (i64, [C.c_char_p, vp, i64]),
            "orc_bench_chains": (d, [vp, i64, vp, vp, i64, C.c_int, i64, C.c_int, d, PP, u64, C.c_int, vp, vp]),
            "orc_build_info": (C.c_char_p, []),
            "orc_go_pow": (d, [d, d]),
            "orc_go_frexp": (d, [d, C.POINTER(C.c_int)]),
            "orc_go_ldexp": (d, [d, C.c_int]),
            "orc_go_modf": (d, [d, C.POINTER(d)]),
        }
        for name, (res, args) in sig.items():
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args

    # ---- params / streams
    def params(self, prot_lj=None, lig_lj=None, **kw) -> OrcParams:
        """prot_lj / lig_lj: (n, 2) arrays of (sigma, epsilon) for the LJ extension; kept alive on the returned struct"""
        p = OrcParams()
        self.L.orc_params_default(C.byref(p))
        if prot_lj is not None:
            p._keep = (np.ascontiguousarray(prot_lj, dtype=np.float64), np.ascontiguousarray(lig_lj, dtype=np.float64))
            p.prot_lj, p.lig_lj = p._keep[0].ctypes.data, p._keep[1].ctypes.data
        for k, v in kw.items():
            if not hasattr(p, k):
                raise AttributeError(k)
            setattr(p, k, v)
        return p

    def philox_stream(self, seed, stream_id) -> OrcStream:
        s = OrcStream()
        self.L.orc_stream_philox(C.byref(s), seed, stream_id)
        return s

    def recorded_stream(self, u: np.ndarray) -> OrcStream:
        s = OrcStream()
        s._keep = np.ascontiguousarray(u, dtype=np.float64)
        self.L.orc_stream_recorded(C.byref(s), _vp(s._keep), len(s._keep))
        return s

    def next(self, s: OrcStream) -> float:
        return self.L.orc_stream_next(C.byref(s))

    def philox(self, ctr, key):
        c = np.array(ctr, dtype=np.uint32); k = np.array(key, dtype=np.uint32); o = np.zeros(4, dtype=np.uint32)
        self.L.orc_philox4x32_10(_vp(c), _vp(k), _vp(o))
        return [int(x) for x in o]

    def u01(self, seed, stream_id, index) -> float:
        return self.L.orc_philox_u01(seed, stream_id, index)

    # ---- vector helpers
    def distance(self, a, b) -> float:
        a = np.array(a, dtype=np.float64); b = np.array(b, dtype=np.float64)
        return self.L.orc_distance(_vp(a), _vp(b))

    def rotate_atom(self, pos, axis, theta):
        p = np.array(pos, dtype=np.float64); a = np.array(axis, dtype=np.float64); o = np.zeros(3)
        self.L.orc_rotate_atom(_vp(p), _vp(a), float(theta), _vp(o))
        return o

    def normalize(self, v):
        v = np.array(v, dtype=np.float64)
        self.L.orc_normalize(_vp(v))
        return v

    # ---- metropolis.go
    def is_collision_free(self, lig, min_distance=0.5) -> bool:
        lig = _atoms(lig)
        return bool(self.L.orc_is_collision_free(_vp(lig), len(lig), float(min_distance)))

    def energy(self, prot, lig, params=None, with_abs=False):
        prot, lig = _atoms(prot), _atoms(lig)
        p = params or self.params()
        if with_abs:
            ab = C.c_double()
            e = self.L.orc_calculate_energy_abs(_vp(prot), len(prot), _vp(lig), len(lig), C.byref(p), C.addressof(ab))
            return e, ab.value
        return self.L.orc_calculate_energy(_vp(prot), len(prot), _vp(lig), len(lig), C.byref(p))

    def energy_batch(self, prot, offsets, xyzq, params=None, with_abs=False):
        n = len(offsets) - 1
        out = np.zeros(n); ab = np.zeros(n)
        for i in range(n):
            r = self.energy(prot, xyzq[offsets[i]:offsets[i + 1]], params, with_abs)
            if with_abs:
                out[i], ab[i] = r
            else:
                out[i] = r
        return (out, ab) if with_abs else out

    def accept_move(self, cur, new, temperature, stream) -> bool:
        return bool(self.L.orc_accept_move(cur, new, temperature, C.byref(stream)))

    def rotate_ligand(self, lig, max_angle, stream):
        lig = np.array(_atoms(lig), copy=True)
        self.L.orc_rotate_ligand(_vp(lig), len(lig), float(max_angle), C.byref(stream))
        return lig

    def jitter_ligand(self, lig, stream, params=None):
        lig = np.array(_atoms(lig), copy=True)
        rej = self.L.orc_jitter_ligand(_vp(lig), len(lig), C.byref(params or self.params()), C.byref(stream))
        return lig, rej

    def jitter_and_rotate_ligand(self, lig, stream, params=None):
        lig = np.array(_atoms(lig), copy=True)
        rej = self.L.orc_jitter_and_rotate_ligand(_vp(lig), len(lig), C.byref(params or self.params()), C.byref(stream))
        return lig, rej

    def rigid_move(self, lig, stream, params=None):
        lig = np.array(_atoms(lig), copy=True)
        self.L.orc_rigid_move(_vp(lig), len(lig), C.byref(params or self.params()), C.byref(stream))
        return lig

    def find_closest(self, lig, prot):
        lig, prot = _atoms(lig), _atoms(prot)
        li, pj = C.c_int64(), C.c_int64()
        d = self.L.orc_find_closest_atom_distance(_vp(lig), len(lig), _vp(prot), len(prot), C.addressof(li), C.addressof(pj))
        return d, li.value, pj.value

    def shift_closer(self, lig, prot, threshold=5.0):
        lig = np.array(_atoms(lig), copy=True); prot = _atoms(prot)
        self.L.orc_shift_ligand_closer(_vp(lig), len(lig), _vp(prot), len(prot), float(threshold))
        return lig

    def chain(self, prot, lig, iterations, rotate, temperature, stream, params=None, want_trace=False):
        """SimulateEnergyMinimizationOneProc on a private copy -> (pose, stats, accept_trace, energy_trace)."""
        prot = _atoms(prot); lig = np.array(_atoms(lig), copy=True)
        st = np.zeros(1, dtype=STATS_DTYPE)
        tr = np.zeros(max(iterations, 1), dtype=np.uint8) if want_trace else None
        et = np.zeros(max(iterations, 1)) if want_trace else None
        self.L.orc_chain_one_proc(_vp(prot), len(prot), _vp(lig), len(lig), iterations, int(bool(rotate)),
                                  float(temperature), C.byref(params or self.params()), C.byref(stream),
                                  _vp(st), _vp(tr), _vp(et))
        if want_trace:
            return lig, st[0], tr[:iterations], et[:iterations]
        return lig, st[0]

    def chain_best(self, prot, lig, iterations, rotate, temperature, stream, params=None):
        """EXTENSION: chain + best-ever pose -> (final pose, stats, best pose, best energy)"""
        prot = _atoms(prot); lig = np.array(_atoms(lig), copy=True)
        st = np.zeros(1, dtype=STATS_DTYPE)
        best = np.zeros_like(lig); be = np.zeros(1)
        self.L.orc_chain_one_proc_best(_vp(prot), len(prot), _vp(lig), len(lig), iterations, int(bool(rotate)),
                                       float(temperature), C.byref(params or self.params()), C.byref(stream),
                                       _vp(st), _vp(best), _vp(be))
        return lig, st[0], best, float(be[0])

    def min_energy_index(self, energies, rule=0):
        """SaveMinimumEnergyLigand (plotting.go:109-121, rule 0) / RShinyAppMain (main.go:51-54, rule 1) -> (index, energy)"""
        e = np.ascontiguousarray(energies, dtype=np.float64)
        out = np.zeros(1)
        i = self.L.orc_min_energy_index(_vp(e), len(e), rule, _vp(out))
        return int(i), float(out[0])

    def minimize_parallel(self, prot, lig, iterations, rotate, temperature, num_procs, seed, lig_index=0, params=None):
        prot = _atoms(prot); lig = np.array(_atoms(lig), copy=True)
        st = np.zeros(max(num_procs, 1), dtype=STATS_DTYPE)
        poses = np.zeros((max(num_procs, 1) * len(lig), 4))
        picked = C.c_int64(-2)
        rc = self.L.orc_minimize_parallel(_vp(prot), len(prot), _vp(lig), len(lig), iterations, int(bool(rotate)),
                                          float(temperature), num_procs, C.byref(params or self.params()), seed,
                                          lig_index, _vp(st), _vp(poses), C.addressof(picked))
        return rc, lig, st, poses, picked.value

    def simulate_multiple_ligands(self, prot, offsets, xyzq, iterations, rotate, temperature, num_procs, seed,
                                  shift_first=False, n_threads=4, params=None):
        prot = _atoms(prot); xyzq = np.array(_atoms(xyzq), copy=True)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        n = len(offsets) - 1
        en = np.zeros(n)
        st = np.zeros(n * max(num_procs, 1), dtype=STATS_DTYPE)
        rc = self.L.orc_simulate_multiple_ligands(_vp(prot), len(prot), _vp(xyzq), _vp(offsets), n, iterations,
                                                  int(bool(rotate)), float(temperature), num_procs,
                                                  C.byref(params or self.params()), seed, int(shift_first),
                                                  n_threads, _vp(en), _vp(st))
        return rc, xyzq, en, st

    def ligand_block(self, n_ligands, num_procs, worker):
        b, e = C.c_int64(), C.c_int64()
        self.L.orc_ligand_block(n_ligands, num_procs, worker, C.addressof(b), C.addressof(e))
        return b.value, e.value

    # ---- validateResults.go
    def rmsd(self, sim, ref) -> float:
        sim, ref = _atoms(sim), _atoms(ref)
        return self.L.orc_calculate_rmsd(_vp(sim), len(sim), _vp(ref), len(ref))

    def randomize_pose(self, lig, stream):
        lig = np.array(_atoms(lig), copy=True)
        self.L.orc_randomize_ligand_pose(_vp(lig), len(lig), C.byref(stream))
        return lig

    def compare_rmsd(self, prot, lig, iterations, rotate, temperature, num_procs, seed, lig_index=0, params=None):
        prot = _atoms(prot); lig = np.array(_atoms(lig), copy=True)
        return self.L.orc_compare_rmsd(_vp(prot), len(prot), _vp(lig), len(lig), iterations, int(bool(rotate)),
                                       float(temperature), num_procs, C.byref(params or self.params()), seed, lig_index)

    # ---- parsing.go
    def parse_mol2(self, path) -> np.ndarray:
        n = self.L.orc_parse_mol2(str(path).encode(), None, 0)
        if n < 0:
            raise OSError(f"cannot open {path}")
        a = np.zeros((n, 4))
        self.L.orc_parse_mol2(str(path).encode(), _vp(a), n)
        return a

    # ---- CPU baseline
    def bench_chains(self, prot, offsets, xyzq, replicas, steps, rotate, temperature, seed, n_threads, params=None):
        prot, xyzq = _atoms(prot), _atoms(xyzq)
        offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        ts, tp = C.c_int64(), C.c_int64()
        sec = self.L.orc_bench_chains(_vp(prot), len(prot), _vp(xyzq), _vp(offsets), len(offsets) - 1, replicas,
                                      steps, int(bool(rotate)), float(temperature), C.byref(params or self.params()),
                                      seed, n_threads, C.addressof(ts), C.addressof(tp))
        return sec, ts.value, tp.value

    def build_info(self) -> str:
        return self.L.orc_build_info().decode()
