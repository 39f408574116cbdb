"""Pins the CPU oracle against every known answer the reference's own tests hold for the hot path
(metropolisMethod/metropolis_test.go).  Each test names the Go test it restates."""
import math

import numpy as np
import pytest

from conftest import mock_ligand, mock_protein


def test_distance_exact(orc):
    # TestDistance (metropolis_test.go:115-124): exact equality with math.Sqrt(50)
    assert orc.distance([1, 2, 3], [4, 6, 8]) == math.sqrt(50)


def test_find_closest_atom_distance_exact(orc):
    # TestFindClosestAtomDistance (:140-148): exact equality with math.Sqrt(12.0)
    d, li, pj = orc.find_closest(mock_ligand(), mock_protein(2.0, 4.0))
    assert d == math.sqrt(12.0)
    assert (li, pj) == (1, 1)          # ligand (3,3,3) vs protein (1,1,1)


def test_rotate_atom(orc):
    # TestRotateAtom (:126-138): (1,0,0) about z by pi/2 -> (0,1,0) within 1e-6
    out = orc.rotate_atom([1, 0, 0], [0, 0, 1], math.pi / 2)
    assert np.allclose(out, [0, 1, 0], atol=1e-6, rtol=0)


def test_calculate_energy_signs_and_values(orc):
    # TestCalculateEnergyPositive (:36-45) / Negative (:47-56); magnitudes derived in SURVEY.md section 4
    pos = orc.energy(mock_protein(1.0, 1.0), mock_ligand(1.0, 1.0))
    neg = orc.energy(mock_protein(1.0, 1.0), mock_ligand(-1.0, -1.0))
    assert pos > 0 and neg < 0
    assert pos == 7361215932.1677294 and neg == -7361215932.1677294
    assert orc.energy(mock_protein(1.0, -1.0), mock_ligand(1.0, -1.0)) == 433012701.89222002


def test_accept_move_downhill(orc):
    # TestAcceptMove (:58-67): new < current -> true, and no draw is consumed (metropolis.go:142-144)
    s = orc.philox_stream(1, 0)
    assert orc.accept_move(10.0, 5.0, 300.0, s)
    assert s.pos == 0


def test_accept_move_uphill_draws_once(orc):
    # the untested branch (metropolis.go:145-148): one draw, compared with exp(-dE/T); dE == 0 -> p = 1
    s = orc.recorded_stream(np.array([0.999999, 0.5, 0.1]))
    assert orc.accept_move(5.0, 5.0, 300.0, s) and s.pos == 1          # u < exp(0) = 1
    assert not orc.accept_move(5.0, 305.0, 300.0, s) and s.pos == 2    # 0.5 < exp(-1) = 0.3679 is false
    assert orc.accept_move(5.0, 305.0, 300.0, s) and s.pos == 3        # 0.1 < 0.3679


def test_jitter_ligand_collision_free(orc):
    # TestJitterLigand (:69-77)
    lig, rej = orc.jitter_ligand(mock_ligand(), orc.philox_stream(3, 0))
    assert rej == 0 and orc.is_collision_free(lig, 0.5)
    assert np.all(np.abs(lig[:, :3] - mock_ligand()[:, :3]) <= 0.05)    # (U-0.5)*0.1
    assert np.array_equal(lig[:, 3], mock_ligand()[:, 3])


def test_rotate_ligand_moves_atom0(orc):
    # TestRotateLigand (:79-88)
    rot = orc.rotate_ligand(mock_ligand(), math.pi / 4, orc.philox_stream(4, 0))
    assert not np.array_equal(rot[0, :3], mock_ligand()[0, :3])
    # a rotation about the ORIGIN preserves |r|
    assert np.allclose(np.linalg.norm(rot[:, :3], axis=1), np.linalg.norm(mock_ligand()[:, :3], axis=1), rtol=1e-14)


def test_shift_ligand_closer_by_threshold(orc):
    # TestShiftLigandCloserByThreshold (:90-101): mock closest = sqrt(12) < 5 -> no shift happens
    shifted = orc.shift_closer(mock_ligand(), mock_protein(1.0, -1.0), 5.0)
    assert np.array_equal(shifted, mock_ligand())
    d1 = orc.distance(shifted[0, :3], [0, 0, 0]); d2 = orc.distance(shifted[1, :3], [1, 1, 1])
    assert not (d1 > 5.0 and d2 > 5.0)
    # and when it is farther than the threshold the closest pair ends up exactly threshold apart
    far = mock_ligand(); far[:, :3] += 10.0
    sh = orc.shift_closer(far, mock_protein(1.0, -1.0), 5.0)
    assert orc.find_closest(sh, mock_protein(1.0, -1.0))[0] == pytest.approx(5.0, abs=1e-12)


def test_copy_semantics_chain_leaves_input_untouched(orc):
    # TestCopyLigand (:103-113): the chain works on a private deep copy
    lig = mock_ligand(2.3, -8.7)
    before = lig.copy()
    orc.chain(mock_protein(1.0, -1.0), lig, 50, True, 300.0, orc.philox_stream(1, 0))
    assert np.array_equal(lig, before)


def test_simulate_multiple_ligands_shapes(orc):
    # TestSimulateMultipleLigands (:8-20): 3 ligands, 1000 it, T=300, numProcs=4 -> 3 outputs
    prot = mock_protein(1.0, -1.0)
    lig = np.concatenate([mock_ligand()] * 3)
    off = np.array([0, 2, 4, 6])
    rc, out, en, st = orc.simulate_multiple_ligands(prot, off, lig, 1000, True, 300.0, 4, seed=1, shift_first=True)
    assert rc == 0 and out.shape == lig.shape and en.shape == (3,)
    assert np.all(st["steps"] == 250)                                   # 1000 / 4


def test_simulate_multiple_ligands_parallel_shapes(orc):
    # TestSimulateMultipleLigandsParallel (:22-34): 5 ligands, 2000 it, numProcs=2 (uneven split 2+3)
    prot = mock_protein(1.0, -1.0)
    lig = np.concatenate([mock_ligand()] * 5)
    off = np.arange(0, 11, 2)
    rc, out, en, st = orc.simulate_multiple_ligands(prot, off, lig, 2000, True, 300.0, 2, seed=1)
    assert rc == 0 and out.shape == lig.shape and en.shape == (5,)
    assert [orc.ligand_block(5, 2, w) for w in range(2)] == [(0, 2), (2, 5)]   # metropolis.go:34-42
    for i in range(5):                                                  # energy[i] = CalculateEnergy(final pose) (:61)
        assert en[i] == orc.energy(prot, out[off[i]:off[i + 1]])


def test_ligand_block_quirk_more_workers_than_ligands(orc):
    # metropolis.go:34: width = 5/8 = 0 -> workers 0..6 empty, the last takes everything
    blocks = [orc.ligand_block(5, 8, w) for w in range(8)]
    assert blocks[:7] == [(0, 0)] * 7 and blocks[7] == (0, 5)


def test_num_procs_zero_is_an_error(orc):
    # Go: iterations / numProcs panics with integer divide by zero (metropolis.go:97)
    rc, *_ = orc.minimize_parallel(mock_protein(1, 1), mock_ligand(), 10, True, 300.0, 0, seed=1)
    assert rc != 0


# ---- Go standard library math.Pow (metropolis.go:253 calls it; third-party to the reference, restated in oracle/)
def test_go_pow_square_is_one_rounded_multiply(orc):
    # math.Pow(x, 2) == x*x bit for bit (frexp, one rounded multiply of the mantissas, ldexp)
    import ctypes as C
    import math
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.normal(0, 40, 4000), rng.uniform(-1e-3, 1e-3, 1000), 10.0 ** rng.uniform(-150, 150, 1000),
                         [0.5, -0.5, 1.0, -1.0, 3.0, 1e-160, 1.5e154, 7.0e-155]])
    for x in xs:
        assert orc.L.orc_go_pow(float(x), 2.0) == float(x) * float(x), x
    # documented special cases of pow.go and a few general values against libm (<= 1 ulp)
    P = orc.L.orc_go_pow
    assert P(5.0, 0.0) == 1.0 and P(1.0, float("nan")) == 1.0 and P(3.0, 1.0) == 3.0
    assert math.isnan(P(float("nan"), 2.0)) and math.isnan(P(-8.0, 1.0 / 3.0))
    assert P(0.0, -1.0) == float("inf") and P(-0.0, -3.0) == float("-inf") and P(-0.0, 3.0) == 0.0
    assert P(2.0, float("inf")) == float("inf") and P(0.5, float("inf")) == 0.0 and P(-1.0, float("inf")) == 1.0
    assert P(float("inf"), -2.0) == 0.0 and P(float("-inf"), 3.0) == float("-inf")
    assert P(2.0, 0.5) == math.sqrt(2.0) and P(4.0, -0.5) == 0.5
    assert P(2.0, 10.0) == 1024.0 and P(2.0, -2.0) == 0.25 and P(-3.0, 3.0) == -27.0
    for x, y in [(1.7, 2.5), (9.1, -1.25), (0.3, 7.0), (123.456, 0.75)]:
        assert P(x, y) == pytest.approx(x ** y, rel=4e-16)
    e = C.c_int()
    assert orc.L.orc_go_frexp(8.0, C.byref(e)) == 0.5 and e.value == 4
    assert orc.L.orc_go_frexp(5e-324, C.byref(e)) == 0.5 and e.value == -1073
    assert orc.L.orc_go_ldexp(0.75, 3) == 6.0 and orc.L.orc_go_ldexp(0.5, -1073) == 5e-324
    fr = C.c_double()
    assert orc.L.orc_go_modf(-3.25, C.byref(fr)) == -3.0 and fr.value == -0.25


def test_gopow_build_is_bit_identical_to_the_oracle():
    # the -DORC_SQUARE_VIA_GOPOW build (the arithmetic Go executes) is a COST model only: same bits as x*x
    from _oracle import Oracle
    a, b = Oracle(), Oracle("gopow")
    prot = np.array([[0.0, 0, 0, 0.3], [1.5, -2, 0.25, -0.41], [7.0, 3, -1, 0.2]])
    lig = np.array([[4.0, 4, 4, 1.0], [3.0, 3, 3, -1.0], [2.5, 3.5, 4.25, 0.17]])
    assert a.energy(prot, lig) == b.energy(prot, lig)
    pa, sa = a.chain(prot, lig, 200, True, 310.15, a.philox_stream(3, 0))
    pb, sb = b.chain(prot, lig, 200, True, 310.15, b.philox_stream(3, 0))
    assert np.array_equal(pa, pb) and sa == sb
    assert "Go math.Pow" in b.build_info()
