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
