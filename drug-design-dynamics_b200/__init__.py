"""drug-design-dynamics_b200 -- B200-native Metropolis Monte-Carlo binding-energy hot path.

Holds only what the hot path of Simran-Sodhi/drug-design-dynamics (metropolisMethod/) needs:

  csrc/        sm_100a CUDA kernels + the C ABI (include/ddd_mc.h) -> csrc/libddd_mc.so
  build.py     nvcc recipe (in-tree build; the .so travels with the snapshot)
  capi.py      ctypes binding of the C ABI (what the cgo overlay in go/ binds, for Python callers)
  metropolis.py  host-side mirror of the reference's Go API for this path (same names/arguments)
  synth.py     synthetic receptor / ligand generators of SURVEY.md 8(d)

The directory name contains a hyphen (it follows the reference repo's name), so it is loaded
through `ddd_loader.load()` at the repo root under the module name `ddd_b200`.
"""
from . import capi, metropolis, synth, build  # noqa: F401

__all__ = ["capi", "metropolis", "synth", "build"]
