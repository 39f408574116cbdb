"""Shared fixtures.  `-m "not gpu"` runs the oracle / host-logic / ABI-export tests on CPU;
`-m gpu` runs the parity tests proper, calling the CUDA library through the C ABI."""
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import ddd_loader  # noqa: E402
from _oracle import Oracle  # noqa: E402

GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def ddd():
    return ddd_loader.load()


@pytest.fixture(scope="session")
def orc():
    return Oracle()


@pytest.fixture(scope="session")
def capi(ddd):
    return ddd.capi


@pytest.fixture(scope="session")
def gpu(ddd):
    """Initialise the CUDA library on device 0; fail loudly (never skip) when that is impossible."""
    n = ddd.capi.init([0])
    assert n == 1
    yield ddd.capi
    ddd.capi.shutdown()


@pytest.fixture(scope="session")
def sample():
    """Shipped sample data (223l/227l proteins + 8 ligands) as xyzq arrays; see golden/make_golden.py."""
    z = np.load(GOLDEN / "sample_mol2.npz")
    return {k: z[k] for k in z.files if not k.endswith("_types")}


@pytest.fixture(scope="session")
def sample_types():
    """SYBYL atom types of the same atoms (LJ extension), keyed like `sample`."""
    z = np.load(GOLDEN / "sample_mol2.npz")
    return {k[:-6]: z[k] for k in z.files if k.endswith("_types")}


@pytest.fixture(scope="session")
def synth_small(ddd):
    prot = ddd.synth.make_receptor(600, seed=11)
    off, lig = ddd.synth.make_ligands([12, 7, 20, 1, 33, 2, 16, 40], prot)
    return prot, off, lig


def mock_protein(c1, c2):
    """metropolis_test.go:153-160 createMockProtein"""
    return np.array([[0, 0, 0, c1], [1, 1, 1, c2]], dtype=np.float64)


def mock_ligand(c1=1.0, c2=-1.0):
    """metropolis_test.go:176-195 createMockLigand / createMockLigandWithCustomCharge"""
    return np.array([[4, 4, 4, c1], [3, 3, 3, c2]], dtype=np.float64)
