"""ctypes binding of the C ABI in include/ddd_mc.h (libddd_mc.so).

This is the Python twin of the cgo stub in go/metropolis_cuda.go: plain pointers and sizes, numpy
arrays as host buffers.  There is no CPU fallback: if the library is missing it is built with nvcc,
and if no CUDA device can be initialised every compute call raises DddError.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import build as _build

# every symbol include/ddd_mc.h declares (tests check that the .so exports exactly these)
SYMBOLS = [
    "ddd_init", "ddd_shutdown", "ddd_num_devices", "ddd_last_error", "ddd_version", "ddd_params_default",
    "ddd_receptor_create", "ddd_receptor_destroy", "ddd_receptor_num_atoms",
    "ddd_energy_batch", "ddd_run_chains", "ddd_minimize_batch", "ddd_rmsd_batch",
    "ddd_closest_atom_batch", "ddd_shift_closer_batch", "ddd_randomize_pose_batch",
    "ddd_screen_create", "ddd_screen_run", "ddd_screen_sync", "ddd_screen_last_kernel_ms",
    "ddd_screen_download", "ddd_screen_destroy", "ddd_screen_set_chain_threads", "ddd_screen_chain_threads",
    "ddd_screen_set_chain_slots", "ddd_screen_chain_slots", "ddd_screen_set_segment_steps", "ddd_screen_segment_steps", "ddd_screen_rmsd", "ddd_screen_pair_units", "ddd_screen_set_lj", "ddd_receptor_set_lj",
    "ddd_shard_plan", "ddd_device_info",
    "ddd_chain_kernel_info",
    "ddd_receptor_acquire", "ddd_receptor_release", "ddd_last_chain_status",
    "ddd_screen_fused_rmsd", "ddd_screen_randomize", "ddd_screen_shift_closer", "ddd_screen_download_start",
    "ddd_screen_pick", "ddd_screen_argmin", "ddd_screen_set_track_best", "ddd_screen_download_best",
]

DDD_OK, DDD_EINVAL, DDD_ENODEVICE, DDD_ECUDA, DDD_ELIMIT = 0, -1, -2, -3, -4
PRECISION_F32, PRECISION_F64 = 0, 1
ARGMIN_SAVE_MIN_LIGAND, ARGMIN_RSHINY = 0, 1
MOVE_REFERENCE, MOVE_RIGID = 0, 1
STATUS_RETRY_LIMIT, STATUS_STREAM_EXHAUSTED = 1, 2
STREAM_PARENT_BASE = 0x4000000000000000
STREAM_RANDOMIZE_BASE = 0x2000000000000000
MAX_LIGAND_ATOMS = 2048


class DddError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"ddd_mc error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    """ddd_params (datatypes.go:6-12 + the literals of metropolis.go)."""
    _fields_ = [
        ("k_coulomb", C.c_double), ("threshold", C.c_double), ("min_distance", C.c_double),
        ("max_angle", C.c_double), ("jitter_width", C.c_double), ("clamp_distance", C.c_double),
        ("rigid_translation", C.c_double), ("rigid_angle", C.c_double),
        ("max_retries", C.c_int64), ("precision", C.c_int32), ("move_mode", C.c_int32),
        ("cutoff", C.c_double), ("lj_scale", C.c_double), ("lj_cap", C.c_double),
    ]


class ChainStats(C.Structure):
    _fields_ = [
        ("steps", C.c_int64), ("accepted", C.c_int64), ("downhill", C.c_int64), ("retries", C.c_int64),
        ("draws", C.c_int64), ("status", C.c_int64),
        ("final_energy", C.c_double), ("start_energy", C.c_double), ("chain_energy", C.c_double),
    ]


STATS_DTYPE = np.dtype([
    ("steps", "<i8"), ("accepted", "<i8"), ("downhill", "<i8"), ("retries", "<i8"), ("draws", "<i8"),
    ("status", "<i8"), ("final_energy", "<f8"), ("start_energy", "<f8"), ("chain_energy", "<f8"),
])
assert STATS_DTYPE.itemsize == C.sizeof(ChainStats)

_lib = None


def library_path() -> Path:
    return _build.LIB


def lib() -> C.CDLL:
    """Load (building first if stale) libddd_mc.so.  Raises if it cannot be built or loaded."""
    global _lib
    if _lib is not None:
        return _lib
    import os
    path = os.environ.get("DDD_MC_LIB") or _build.build_library()     # DDD_MC_LIB: tuning builds (tools/), never a fallback
    L = C.CDLL(str(path))
    vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
    sig = {
        "ddd_init": (C.c_int, [vp, C.c_int]),
        "ddd_shutdown": (C.c_int, []),
        "ddd_num_devices": (C.c_int, []),
        "ddd_last_error": (C.c_char_p, []),
        "ddd_version": (C.c_char_p, []),
        "ddd_params_default": (None, [C.POINTER(Params)]),
        "ddd_receptor_create": (C.c_int, [vp, i64, C.POINTER(vp)]),
        "ddd_receptor_destroy": (C.c_int, [vp]),
        "ddd_receptor_num_atoms": (i64, [vp]),
        "ddd_energy_batch": (C.c_int, [vp, vp, vp, i64, C.POINTER(Params), vp, vp]),
        "ddd_run_chains": (C.c_int, [vp, vp, vp, i64, i32, i64, i32, dbl, C.POINTER(Params), u64, u64,
                                     vp, vp, vp, vp, vp]),
        "ddd_minimize_batch": (C.c_int, [vp, vp, vp, i64, i64, i32, dbl, i32, C.POINTER(Params), u64, u64,
                                         vp, vp, vp, vp]),
        "ddd_rmsd_batch": (C.c_int, [vp, vp, vp, i32, i64, vp]),
        "ddd_closest_atom_batch": (C.c_int, [vp, vp, vp, i64, vp, vp, vp]),
        "ddd_shift_closer_batch": (C.c_int, [vp, vp, vp, i64, dbl]),
        "ddd_randomize_pose_batch": (C.c_int, [vp, vp, i64, u64, u64]),
        "ddd_screen_create": (C.c_int, [vp, vp, vp, i64, i32, u64, C.POINTER(vp)]),
        "ddd_screen_run": (C.c_int, [vp, i64, i32, dbl, C.POINTER(Params), u64]),
        "ddd_screen_sync": (C.c_int, [vp]),
        "ddd_screen_set_chain_threads": (C.c_int, [vp, i32]),
        "ddd_screen_chain_threads": (i32, [vp]),
        "ddd_screen_set_chain_slots": (C.c_int, [vp, i32]),
        "ddd_screen_chain_slots": (i32, [vp]),
        "ddd_screen_set_segment_steps": (C.c_int, [vp, i64]),
        "ddd_screen_segment_steps": (i64, [vp]),
        "ddd_screen_rmsd": (C.c_int, [vp, vp, C.POINTER(dbl)]),
        "ddd_screen_pair_units": (C.c_int, [vp, vp]),
        "ddd_screen_set_lj": (C.c_int, [vp, vp, vp]),
        "ddd_receptor_set_lj": (C.c_int, [vp, vp, vp]),
        "ddd_screen_last_kernel_ms": (dbl, [vp]),
        "ddd_screen_download": (C.c_int, [vp, vp, vp]),
        "ddd_screen_destroy": (C.c_int, [vp]),
        "ddd_shard_plan": (C.c_int, [vp, i64, i32, i32, vp]),
        "ddd_device_info": (C.c_int, [C.c_int, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32),
                                      C.POINTER(i32), C.c_char_p, i32]),
        "ddd_chain_kernel_info": (C.c_int, [i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]),
        "ddd_receptor_acquire": (C.c_int, [vp, i64, C.POINTER(vp)]),
        "ddd_receptor_release": (C.c_int, [vp]),
        "ddd_last_chain_status": (i64, [C.POINTER(i64)]),
        "ddd_screen_fused_rmsd": (C.c_int, [vp, vp]),
        "ddd_screen_randomize": (C.c_int, [vp, u64, u64]),
        "ddd_screen_shift_closer": (C.c_int, [vp, dbl]),
        "ddd_screen_download_start": (C.c_int, [vp, vp]),
        "ddd_screen_pick": (C.c_int, [vp, dbl, u64, u64, vp, vp, vp, vp]),
        "ddd_screen_argmin": (C.c_int, [vp, i32, C.POINTER(i64), C.POINTER(dbl)]),
        "ddd_screen_set_track_best": (C.c_int, [vp, i32]),
        "ddd_screen_download_best": (C.c_int, [vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _check(rc: int) -> None:
    if rc != DDD_OK:
        raise DddError(rc, lib().ddd_last_error().decode(errors="replace"))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, shape_last=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape_last is not None and a.size and a.shape[-1] != shape_last:
        raise ValueError(f"expected last dimension {shape_last}, got {a.shape}")
    return a


def _i64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int64)


def default_params(**overrides) -> Params:
    p = Params()
    lib().ddd_params_default(C.byref(p))
    for k, v in overrides.items():
        if not hasattr(p, k):
            raise AttributeError(k)
        setattr(p, k, v)
    return p


def init(devices=None) -> int:
    """ddd_init.  devices: list of CUDA ordinals, or None for all visible devices."""
    if devices is None:
        _check(lib().ddd_init(None, 0))
    else:
        arr = (C.c_int * len(devices))(*devices)
        _check(lib().ddd_init(arr, len(devices)))
    return lib().ddd_num_devices()


def shutdown() -> None:
    _check(lib().ddd_shutdown())


def num_devices() -> int:
    return lib().ddd_num_devices()


def version() -> str:
    return lib().ddd_version().decode()


def pack_ligands(ligands):
    """list of (nL,4) xyzq arrays -> (offsets int64[n+1], xyzq float64[sum nL, 4])."""
    offsets = np.zeros(len(ligands) + 1, dtype=np.int64)
    for i, l in enumerate(ligands):
        offsets[i + 1] = offsets[i] + len(l)
    if len(ligands) and offsets[-1] > 0:
        xyzq = np.concatenate([_f64(l).reshape(-1, 4) for l in ligands], axis=0)
    else:
        xyzq = np.zeros((0, 4))
    return offsets, np.ascontiguousarray(xyzq)


def unpack_ligands(offsets, xyzq):
    return [np.array(xyzq[offsets[i]:offsets[i + 1]]) for i in range(len(offsets) - 1)]


class Receptor:
    """ddd_receptor: the protein Molecule (datatypes.go:14-16) resident on every device."""

    def __init__(self, xyzq, cached=False):
        """cached=True: ddd_receptor_acquire -- the library's content-keyed cache (what the scalar API mirrors use)."""
        a = _f64(xyzq).reshape(-1, 4)
        self.n_atoms = a.shape[0]
        self._h = C.c_void_p()
        self._cached = cached
        if cached:
            _check(lib().ddd_receptor_acquire(_ptr(a), a.shape[0], C.byref(self._h)))
        else:
            _check(lib().ddd_receptor_create(_ptr(a), a.shape[0], C.byref(self._h)))

    @property
    def handle(self):
        if not self._h:
            raise DddError(DDD_EINVAL, "receptor already destroyed")
        return self._h

    def set_lj(self, sigma, epsilon):
        """LJ extension: per-atom sigma / epsilon of the receptor"""
        sg, ep = np.ascontiguousarray(sigma, dtype=np.float64), np.ascontiguousarray(epsilon, dtype=np.float64)
        assert len(sg) == len(ep) == self.n_atoms
        _check(lib().ddd_receptor_set_lj(self.handle, _ptr(sg), _ptr(ep)))

    def close(self):
        if self._h:
            if self._cached:
                lib().ddd_receptor_release(self._h)
            else:
                lib().ddd_receptor_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def energy_batch(rec: Receptor, offsets, xyzq, params: Params | None = None, with_abs=False):
    """CalculateEnergy (metropolis.go:247-262) for every ligand pose of the CSR batch."""
    p = params or default_params()
    offsets, xyzq = _i64(offsets), _f64(xyzq, 4)
    n = len(offsets) - 1
    out = np.zeros(n)
    ab = np.zeros(n) if with_abs else None
    _check(lib().ddd_energy_batch(rec.handle, _ptr(offsets), _ptr(xyzq), n, C.byref(p), _ptr(out), _ptr(ab)))
    return (out, ab) if with_abs else out


def run_chains(rec: Receptor, offsets, xyzq, replicas, steps, rotate, temperature, params: Params | None = None,
               seed=1, first_chain_id=0, recorded=None, recorded_offsets=None, want_trace=False):
    """SimulateEnergyMinimizationOneProc (metropolis.go:116-136) for n_lig*replicas chains.
    Returns (poses[(replicas*sum nL), 4] chain-major, stats structured array, trace or None)."""
    p = params or default_params()
    offsets, xyzq = _i64(offsets), _f64(xyzq, 4)
    n = len(offsets) - 1
    n_chains = n * replicas
    out = np.zeros((int(offsets[-1]) * replicas if n else 0, 4))
    stats = np.zeros(n_chains, dtype=STATS_DTYPE)
    trace = np.zeros((n_chains, steps), dtype=np.uint8) if want_trace else None
    rec_a = _f64(recorded) if recorded is not None else None
    rec_o = _i64(recorded_offsets) if recorded_offsets is not None else None
    _check(lib().ddd_run_chains(rec.handle, _ptr(offsets), _ptr(xyzq), n, replicas, steps, int(bool(rotate)),
                                float(temperature), C.byref(p), seed, first_chain_id, _ptr(rec_a), _ptr(rec_o),
                                _ptr(out), _ptr(stats), _ptr(trace)))
    return out, stats, trace


def chain_pose(offsets, replicas, poses, lig, rep):
    """Slice chain (lig, rep) out of the chain-major pose array returned by run_chains."""
    nl = int(offsets[lig + 1] - offsets[lig])
    start = replicas * int(offsets[lig]) + rep * nl
    return poses[start:start + nl]


def minimize_batch(rec: Receptor, offsets, xyzq, iterations, rotate, temperature, num_procs,
                   params: Params | None = None, seed=1, first_ligand_index=0, want_stats=False):
    """SimulateEnergyMinimizationParallel (:94-111) + final CalculateEnergy (:61) for every ligand."""
    p = params or default_params()
    offsets, xyzq = _i64(offsets), _f64(xyzq, 4)
    n = len(offsets) - 1
    out = np.zeros_like(xyzq)
    energy = np.zeros(n)
    stats = np.zeros(n * max(num_procs, 0), dtype=STATS_DTYPE) if want_stats else None
    picked = np.zeros(n, dtype=np.int64)
    _check(lib().ddd_minimize_batch(rec.handle, _ptr(offsets), _ptr(xyzq), n, iterations, int(bool(rotate)),
                                    float(temperature), num_procs, C.byref(p), seed, first_ligand_index,
                                    _ptr(out), _ptr(energy), _ptr(stats), _ptr(picked)))
    return (out, energy, picked, stats) if want_stats else (out, energy, picked)


def last_chain_status():
    """(OR of the DDD_STATUS_* bits, chains stopped at the retry bound) of this thread's last minimize_batch / run_chains."""
    n = C.c_int64()
    bits = lib().ddd_last_chain_status(C.byref(n))
    return int(bits), int(n.value)


def rmsd_batch(offsets, sim, ref, atom_stride=4):
    """CalculateRMSD (validateResults.go:50-65) for every (simulated, reference) pair."""
    offsets = _i64(offsets)
    sim, ref = _f64(sim, atom_stride), _f64(ref, atom_stride)
    n = len(offsets) - 1
    out = np.zeros(n)
    _check(lib().ddd_rmsd_batch(_ptr(offsets), _ptr(sim), _ptr(ref), atom_stride, n, _ptr(out)))
    return out


def closest_atom_batch(rec: Receptor, offsets, xyzq):
    """FindClosestAtomDistance (:290-304): (distance, ligand atom index, protein atom index) per ligand."""
    offsets, xyzq = _i64(offsets), _f64(xyzq, 4)
    n = len(offsets) - 1
    d = np.zeros(n); li = np.zeros(n, dtype=np.int64); pj = np.zeros(n, dtype=np.int64)
    _check(lib().ddd_closest_atom_batch(rec.handle, _ptr(offsets), _ptr(xyzq), n, _ptr(d), _ptr(li), _ptr(pj)))
    return d, li, pj


def shift_closer_batch(rec: Receptor, offsets, xyzq, threshold=5.0):
    """ShiftLigandCloserByThreshold (:267-285); returns the shifted copy."""
    offsets = _i64(offsets)
    out = np.array(_f64(xyzq, 4), copy=True)
    _check(lib().ddd_shift_closer_batch(rec.handle, _ptr(offsets), _ptr(out), len(offsets) - 1, float(threshold)))
    return out


def randomize_pose_batch(offsets, xyzq, seed=1, first_ligand_index=0):
    """RandomizeLigandPose (validateResults.go:70-79); returns the randomised copy."""
    offsets = _i64(offsets)
    out = np.array(_f64(xyzq, 4), copy=True)
    _check(lib().ddd_randomize_pose_batch(_ptr(offsets), _ptr(out), len(offsets) - 1, seed, first_ligand_index))
    return out


def shard_plan(offsets, replicas, n_shards):
    """Host-only: chain ranges per shard (contiguous, balanced by sum(nL+1)); int64[n_shards+1]."""
    offsets = _i64(offsets)
    out = np.zeros(n_shards + 1, dtype=np.int64)
    _check(lib().ddd_shard_plan(_ptr(offsets), len(offsets) - 1, replicas, n_shards, _ptr(out)))
    return out


class Screen:
    """ddd_screen: chains resident in HBM (bench `value`: inputs already on device)."""

    def __init__(self, rec: Receptor, offsets, xyzq, replicas=1, first_chain_id=0):
        self.offsets, xyzq = _i64(offsets), _f64(xyzq, 4)
        self.replicas = replicas
        self.n_lig = len(self.offsets) - 1
        self.n_chains = self.n_lig * replicas
        self._rec = rec
        self._h = C.c_void_p()
        _check(lib().ddd_screen_create(rec.handle, _ptr(self.offsets), _ptr(xyzq), self.n_lig, replicas,
                                       first_chain_id, C.byref(self._h)))

    def run(self, steps, rotate, temperature, params: Params | None = None, seed=1):
        p = params or default_params()
        _check(lib().ddd_screen_run(self._h, steps, int(bool(rotate)), float(temperature), C.byref(p), seed))

    def set_chain_threads(self, threads: int):
        _check(lib().ddd_screen_set_chain_threads(self._h, threads))

    def chain_threads(self) -> int:
        return int(lib().ddd_screen_chain_threads(self._h))

    def set_chain_slots(self, slots: int):
        """chains resident per SM in the SM-persistent engine (0 = automatic)"""
        _check(lib().ddd_screen_set_chain_slots(self._h, slots))

    def chain_slots(self) -> int:
        return int(lib().ddd_screen_chain_slots(self._h))

    def set_segment_steps(self, steps: int):
        """chain segments of the SM-persistent engine: -1 automatic, 0 whole chains, > 0 steps per segment (power of two)"""
        _check(lib().ddd_screen_set_segment_steps(self._h, steps))

    def segment_steps(self) -> int:
        return int(lib().ddd_screen_segment_steps(self._h))

    def sync(self):
        _check(lib().ddd_screen_sync(self._h))

    def last_kernel_ms(self) -> float:
        return float(lib().ddd_screen_last_kernel_ms(self._h))

    def download(self):
        out = np.zeros((int(self.offsets[-1]) * self.replicas if self.n_lig else 0, 4))
        stats = np.zeros(self.n_chains, dtype=STATS_DTYPE)
        _check(lib().ddd_screen_download(self._h, _ptr(out), _ptr(stats)))
        return out, stats

    def set_lj(self, sigma, epsilon):
        """LJ extension: per-atom sigma / epsilon of the ligands (CSR order)"""
        sg, ep = np.ascontiguousarray(sigma, dtype=np.float64), np.ascontiguousarray(epsilon, dtype=np.float64)
        assert len(sg) == len(ep) == int(self.offsets[-1])
        _check(lib().ddd_screen_set_lj(self._h, _ptr(sg), _ptr(ep)))

    def pair_units(self):
        """cutoff extension: receptor atoms visited by each chain's pair sums, summed over its evaluations"""
        out = np.zeros(self.n_chains, dtype=np.int64)
        _check(lib().ddd_screen_pair_units(self._h, _ptr(out)))
        return out

    def rmsd(self):
        """CompareRMSD per chain (final pose vs the ligand's start pose), device-resident; returns (rmsd[n_chains], kernel ms)."""
        out = np.zeros(self.n_chains)
        ms = C.c_double()
        _check(lib().ddd_screen_rmsd(self._h, _ptr(out), C.byref(ms)))
        return out, ms.value

    def fused_rmsd(self):
        """the per-chain RMSD the chain kernel's epilogue computed (SM-persistent engine); same bits as rmsd()"""
        out = np.zeros(self.n_chains)
        _check(lib().ddd_screen_fused_rmsd(self._h, _ptr(out)))
        return out

    # ---- resident pipeline (validateResults.go:14-45, metropolis.go:14-16, plotting.go:109-121)
    def randomize(self, seed=1, first_ligand_index=0):
        _check(lib().ddd_screen_randomize(self._h, seed, first_ligand_index))

    def shift_closer(self, threshold=5.0):
        _check(lib().ddd_screen_shift_closer(self._h, float(threshold)))

    def download_start(self):
        out = np.zeros((int(self.offsets[-1]) if self.n_lig else 0, 4))
        _check(lib().ddd_screen_download_start(self._h, _ptr(out)))
        return out

    def pick(self, temperature, seed=1, first_ligand_index=0, download=True):
        """replica pick on the device -> (pose[sum nL, 4], energy[n_lig], picked[n_lig], rmsd[n_lig]) (or None when download=False)"""
        if not download:
            _check(lib().ddd_screen_pick(self._h, float(temperature), seed, first_ligand_index, None, None, None, None))
            return None
        out = np.zeros((int(self.offsets[-1]) if self.n_lig else 0, 4))
        en = np.zeros(self.n_lig); pk = np.zeros(self.n_lig, dtype=np.int64); rm = np.zeros(self.n_lig)
        _check(lib().ddd_screen_pick(self._h, float(temperature), seed, first_ligand_index, _ptr(out), _ptr(en), _ptr(pk), _ptr(rm)))
        return out, en, pk, rm

    def argmin(self, rule=ARGMIN_SAVE_MIN_LIGAND):
        idx, en = C.c_int64(), C.c_double()
        _check(lib().ddd_screen_argmin(self._h, rule, C.byref(idx), C.byref(en)))
        return int(idx.value), float(en.value)

    def set_track_best(self, on=True):
        _check(lib().ddd_screen_set_track_best(self._h, int(bool(on))))

    def download_best(self):
        out = np.zeros((int(self.offsets[-1]) * self.replicas if self.n_lig else 0, 4))
        en = np.zeros(self.n_chains)
        _check(lib().ddd_screen_download_best(self._h, _ptr(out), _ptr(en)))
        return out, en

    def close(self):
        if self._h:
            lib().ddd_screen_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def device_info(ordinal=0) -> dict:
    sm, khz, mj, mn = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    name = C.create_string_buffer(256)
    _check(lib().ddd_device_info(ordinal, C.byref(sm), C.byref(khz), C.byref(mj), C.byref(mn), name, 256))
    return {"sm_count": sm.value, "clock_khz": khz.value, "cc": (mj.value, mn.value), "name": name.value.decode()}


def chain_kernel_info(precision=PRECISION_F32) -> dict:
    t, r, s, b = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    _check(lib().ddd_chain_kernel_info(precision, C.byref(t), C.byref(r), C.byref(s), C.byref(b)))
    return {"threads": t.value, "regs": r.value, "static_smem": s.value, "blocks_per_sm_nl40": b.value}
