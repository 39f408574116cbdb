"""The C-ABI library builds for sm_100a, loads on a CPU-only box and exports every symbol that
include/ddd_mc.h declares; compute entry points fail loudly without a device (no CPU fallback)."""
import ctypes
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    text = (ROOT / "include" / "ddd_mc.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ddd_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_all_exported(capi):
    lib = capi.lib()
    declared = _declared_symbols()
    assert len(declared) >= 24
    for name in declared:
        assert hasattr(lib, name), f"libddd_mc.so does not export {name}"
    assert sorted(capi.SYMBOLS) == declared


def test_library_contains_sm100a_code_only(capi):
    out = subprocess.run(["cuobjdump", "-lelf", str(capi.library_path())], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_struct_layouts(capi):
    assert ctypes.sizeof(capi.Params) == 104
    assert ctypes.sizeof(capi.ChainStats) == 72 == capi.STATS_DTYPE.itemsize
    p = capi.default_params()
    assert (p.k_coulomb, p.threshold, p.min_distance, p.max_angle, p.jitter_width, p.clamp_distance) == \
        (9e9, 5.0, 0.5, 0.79, 0.1, 1e-6)                  # datatypes.go:6-12, metropolis.go:158,255
    assert p.precision == capi.PRECISION_F32 and p.move_mode == capi.MOVE_REFERENCE and p.max_retries == 1024
    assert p.cutoff == 0.0 and p.lj_scale == 0.0          # extensions are off by default: every pair, Coulomb only


def test_no_device_means_loud_failure_not_fallback(capi):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path is exercised on the CPU box")
    with pytest.raises(capi.DddError) as ei:
        capi.init()
    assert ei.value.code == capi.DDD_ENODEVICE and "no CPU fallback" in str(ei.value)
    with pytest.raises(capi.DddError) as ei:
        capi.Receptor(np.zeros((4, 4)))
    assert ei.value.code == capi.DDD_ENODEVICE
    with pytest.raises(capi.DddError):
        capi.rmsd_batch(np.array([0, 1]), np.zeros((1, 4)), np.zeros((1, 4)))


def test_product_never_references_the_oracle():
    pkg = ROOT / "drug-design-dynamics_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cuh")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.hpp")):
        text = f.read_text()
        if f.name == "build.py":
            text = text.split("def build_oracle")[0]          # building the checker is not using it
        assert "ddd_oracle" not in text and "_oracle" not in text and "orc_" not in text, f
