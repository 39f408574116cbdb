"""-m gpu: CalculateEnergy (metropolis.go:247-262) on device vs the CPU oracle and the golden vectors.

Tolerances:
  F64 kernel: |dE| <= 1e-13 * sum|term|  (same per-term bits as Go; only the summation order differs)
  F32 kernel: |dE| <= 1e-6 * sum|term|, and <= 1e-5 * |E| wherever the sum is well conditioned (cond <= 64)
"""
import json

import numpy as np
import pytest

from conftest import GOLDEN, mock_ligand, mock_protein

pytestmark = pytest.mark.gpu

F64_TOL, F32_ABS_TOL, F32_REL_TOL = 1e-13, 1e-6, 1e-5


def _check(e, eo, ao, f32):
    tiny = 1e-300
    if f32:
        assert np.all(np.abs(e - eo) <= F32_ABS_TOL * ao + tiny)
        well = ao <= 64 * np.abs(eo)
        assert np.all(np.abs(e - eo)[well] <= F32_REL_TOL * np.abs(eo)[well])
    else:
        assert np.all(np.abs(e - eo) <= F64_TOL * ao + tiny)


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_energy_reference_mock_vectors(gpu, precision):
    # TestCalculateEnergyPositive / Negative (metropolis_test.go:36-56) + the derived magnitudes
    p = gpu.default_params(precision=gpu.PRECISION_F64 if precision == "f64" else gpu.PRECISION_F32)
    off = np.array([0, 2])
    want = {(1.0, 1.0, 1.0, 1.0): 7361215932.1677294, (1.0, 1.0, -1.0, -1.0): -7361215932.1677294,
            (1.0, -1.0, 1.0, -1.0): 433012701.89222002}
    for (p1, p2, l1, l2), e_want in want.items():
        with gpu.Receptor(mock_protein(p1, p2)) as rec:
            e = gpu.energy_batch(rec, off, mock_ligand(l1, l2), p)[0]
        assert e == pytest.approx(e_want, rel=1e-15 if precision == "f64" else 2e-6)
        assert np.sign(e) == np.sign(e_want)


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_energy_sample_data_golden(gpu, sample, precision):
    # shipped 223l / 227l proteins x 8 ligands; includes the neutral ligand 7loc with cond 5.5e5 / 1.5e6
    g = json.loads((GOLDEN / "oracle_golden.json").read_text())["sample_energy"]
    p = gpu.default_params(precision=gpu.PRECISION_F64 if precision == "f64" else gpu.PRECISION_F32)
    for prot in ("223l_protein", "227l_protein"):
        names = [k.split("|")[1] for k in g if k.startswith(prot)]
        off, lig = gpu.pack_ligands([sample[n + "_ligand"] for n in names])
        with gpu.Receptor(sample[prot]) as rec:
            e, ab = gpu.energy_batch(rec, off, lig, p, with_abs=True)
        eo = np.array([g[f"{prot}|{n}"]["energy"] for n in names]); ao = np.array([g[f"{prot}|{n}"]["abs_sum"] for n in names])
        _check(e, eo, ao, precision == "f32")
        assert np.allclose(ab, ao, rtol=1e-12)


@pytest.mark.parametrize("precision", ["f64", "f32"])
def test_energy_synthetic_vs_oracle(gpu, orc, ddd, precision):
    prot = ddd.synth.make_receptor(5000)
    off, lig = ddd.synth.make_ligands([40] * 12 + [20, 33, 57, 80, 1, 2, 3, 141], prot)
    eo, ao = orc.energy_batch(prot, off, lig, with_abs=True)
    p = gpu.default_params(precision=gpu.PRECISION_F64 if precision == "f64" else gpu.PRECISION_F32)
    with gpu.Receptor(prot) as rec:
        e = gpu.energy_batch(rec, off, lig, p)
    _check(e, eo, ao, precision == "f32")


def test_energy_receptor_sizes_and_padding(gpu, orc, ddd):
    # receptor sizes around every tile boundary of the fp32 kernel (128 / 256 / 512 atoms), and odd ligands
    prot_all = ddd.synth.make_receptor(1100, seed=3)
    off, lig = ddd.synth.make_ligands([5, 8, 17], prot_all)
    for n in (1, 2, 127, 128, 129, 255, 256, 257, 511, 512, 513, 1024, 1100):
        prot = prot_all[:n]
        eo, ao = orc.energy_batch(prot, off, lig, with_abs=True)
        with gpu.Receptor(prot) as rec:
            _check(gpu.energy_batch(rec, off, lig, gpu.default_params(precision=gpu.PRECISION_F32)), eo, ao, True)
            _check(gpu.energy_batch(rec, off, lig, gpu.default_params(precision=gpu.PRECISION_F64)), eo, ao, False)


def test_energy_clamp_and_coincident_atoms(gpu, orc):
    # metropolis.go:255-257: d < 1e-6 -> 1e-6, also at d == 0 (rsqrt(0) = inf must clamp, not NaN)
    prot = np.array([[1.0, 2.0, 3.0, 0.5], [4.0, 4.0, 4.0, -0.3]])
    lig = np.array([[1.0, 2.0, 3.0, -0.25], [1.0 + 5e-7, 2.0, 3.0, 0.1], [9.0, 9.0, 9.0, 0.2]])
    off = np.array([0, 3])
    eo, ao = orc.energy(prot, lig, with_abs=True)
    with gpu.Receptor(prot) as rec:
        e64 = gpu.energy_batch(rec, off, lig, gpu.default_params(precision=gpu.PRECISION_F64))[0]
        e32 = gpu.energy_batch(rec, off, lig, gpu.default_params(precision=gpu.PRECISION_F32))[0]
    assert e64 == pytest.approx(eo, rel=1e-14)
    assert np.isfinite(e32) and abs(e32 - eo) <= 1e-6 * ao


def test_energy_empty_and_ragged(gpu):
    prot = np.array([[0.0, 0, 0, 1.0]])
    with gpu.Receptor(prot) as rec:
        assert gpu.energy_batch(rec, np.array([0]), np.zeros((0, 4))).shape == (0,)            # no ligands
        e = gpu.energy_batch(rec, np.array([0, 0, 1, 1]), np.array([[3.0, 0, 0, 2.0]]))        # empty ligands in the batch
        assert e[0] == 0.0 and e[2] == 0.0 and e[1] == pytest.approx(9e9 * 2.0 / 3.0, rel=1e-6)
    with gpu.Receptor(np.zeros((0, 4))) as rec:                                                 # empty protein
        assert gpu.energy_batch(rec, np.array([0, 1]), np.array([[3.0, 0, 0, 2.0]]))[0] == 0.0


def test_energy_linearity_in_charge_full_size(gpu, ddd):
    # size-independent property at BASELINE scale: E is bilinear in the charges; doubling ligand charges doubles E,
    # and E(q_a + q_b) = E(q_a) + E(q_b) to fp64 rounding in the F64 kernel
    prot = ddd.synth.make_receptor(5000)
    off, lig = ddd.synth.make_ligands([40] * 256, prot)
    p64 = gpu.default_params(precision=gpu.PRECISION_F64)
    rng = np.random.default_rng(1)
    qa, qb = rng.normal(0, 0.2, len(lig)), rng.normal(0, 0.2, len(lig))
    la, lb, lab = lig.copy(), lig.copy(), lig.copy()
    la[:, 3], lb[:, 3], lab[:, 3] = qa, qb, qa + qb
    with gpu.Receptor(prot) as rec:
        ea, aa = gpu.energy_batch(rec, off, la, p64, with_abs=True)
        eb, ab = gpu.energy_batch(rec, off, lb, p64, with_abs=True)
        eab = gpu.energy_batch(rec, off, lab, p64)
        e2 = gpu.energy_batch(rec, off, np.concatenate([la[:, :3], 2 * la[:, 3:]], axis=1), p64)
        e32 = gpu.energy_batch(rec, off, la, gpu.default_params(precision=gpu.PRECISION_F32))
    assert np.all(np.abs(eab - (ea + eb)) <= 1e-12 * (aa + ab))
    assert np.array_equal(e2, 2 * ea)
    assert np.all(np.abs(e32 - ea) <= F32_ABS_TOL * aa)


def test_energy_translation_invariance(gpu, ddd):
    # moving protein and ligand together must not change E beyond coordinate rounding (fp32 frame is receptor-centred)
    prot = ddd.synth.make_receptor(800, seed=5)
    off, lig = ddd.synth.make_ligands([30, 31], prot)
    t = np.array([512.0, -256.0, 128.0, 0.0])         # exactly representable shift
    with gpu.Receptor(prot) as r0, gpu.Receptor(prot + t) as r1:
        e0, a0 = gpu.energy_batch(r0, off, lig, gpu.default_params(), with_abs=True)
        e1 = gpu.energy_batch(r1, off, lig + t, gpu.default_params())
    assert np.all(np.abs(e0 - e1) <= 2e-6 * a0)
