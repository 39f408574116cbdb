"""-m gpu: batched RMSD (validateResults.go:50-65), FindClosestAtomDistance / ShiftLigandCloserByThreshold
(metropolis.go:267-304), RandomizeLigandPose (validateResults.go:70-79) and the reference-API mirror."""
import json
import math

import numpy as np
import pytest

from conftest import GOLDEN, mock_ligand, mock_protein

pytestmark = pytest.mark.gpu


def test_rmsd_batch(gpu, orc, synth_small):
    prot, off, lig = synth_small
    g = json.loads((GOLDEN / "oracle_golden.json").read_text())["synth600"]
    rz = gpu.randomize_pose_batch(off, lig, seed=5)
    assert np.allclose([float(rz[off[i]:off[i + 1], :3].sum()) for i in range(8)], g["randomize_checksum"], atol=1e-9)
    r4 = gpu.rmsd_batch(off, rz, lig, 4)
    r3 = gpu.rmsd_batch(off, np.ascontiguousarray(rz[:, :3]), np.ascontiguousarray(lig[:, :3]), 3)
    assert np.allclose(r4, g["rmsd"], rtol=1e-13) and np.allclose(r3, g["rmsd"], rtol=1e-13)
    assert np.allclose(gpu.rmsd_batch(off, lig, lig), 0.0)
    # empty pair -> NaN (0/0 as in Go); empty batch -> empty
    r = gpu.rmsd_batch(np.array([0, 0, 2]), mock_ligand(), mock_ligand())
    assert math.isnan(r[0]) and r[1] == 0.0
    assert gpu.rmsd_batch(np.array([0]), np.zeros((0, 4)), np.zeros((0, 4))).shape == (0,)
    with pytest.raises(gpu.DddError):                                        # only Go's []Atom (4) or packed xyz (3)
        gpu.rmsd_batch(off, np.zeros((len(lig), 5)), np.zeros((len(lig), 5)), atom_stride=5)


def test_rmsd_batch_full_size_properties(gpu, ddd):
    # config 5 shape: 100 000 pose pairs, 20-80 atoms; a rigid translation by t gives RMSD = |t| exactly-ish
    sizes = ddd.synth.sweep_sizes(100000)
    off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    rng = np.random.default_rng(0)
    ref = rng.normal(0, 10, size=(off[-1], 4))
    t = rng.normal(0, 3, size=(100000, 3))
    sim = ref.copy()
    sim[:, :3] += np.repeat(t, sizes, axis=0)
    r = gpu.rmsd_batch(off, sim, ref, 4)
    assert np.allclose(r, np.linalg.norm(t, axis=1), rtol=1e-9)
    assert np.array_equal(gpu.rmsd_batch(off, ref, sim, 4), r)               # symmetric


def test_closest_atom_and_shift(gpu, orc, synth_small):
    prot, off, lig = synth_small
    g = json.loads((GOLDEN / "oracle_golden.json").read_text())["synth600"]
    with gpu.Receptor(prot) as rec:
        d, li, pj = gpu.closest_atom_batch(rec, off, lig)
        assert [[d[i], li[i], pj[i]] for i in range(8)] == g["closest"]     # exact distances, first-minimum indices
        far = lig.copy(); far[:, :3] += 25.0
        sh = gpu.shift_closer_batch(rec, off, far, 5.0)
        assert np.allclose([float(sh[off[i]:off[i + 1], :3].sum()) for i in range(8)], g["shift_checksum"], rtol=0, atol=1e-9)
        for i in range(8):
            assert np.array_equal(sh[off[i]:off[i + 1]], orc.shift_closer(far[off[i]:off[i + 1]], prot, 5.0))
        d2, _, _ = gpu.closest_atom_batch(rec, off, sh)
        assert np.all(d2 <= 5.0 + 1e-9)
        assert np.array_equal(gpu.shift_closer_batch(rec, off, lig, 50.0), lig)       # already closer: untouched


def test_closest_ties_pick_first_in_reference_order(gpu, orc):
    # four equidistant pairs: ligand-outer / protein-inner with strict '<' keeps the first (metropolis.go:293-300)
    prot = np.array([[1.0, 0, 0, 1], [-1.0, 0, 0, 1], [0, 1.0, 0, 1]])
    lig = np.array([[0.0, 0, 5.0, 1], [0.0, 0, 0.0, 1], [0.0, 0, 0.0, 1]])
    with gpu.Receptor(prot) as rec:
        d, li, pj = gpu.closest_atom_batch(rec, np.array([0, 3]), lig)
    assert (d[0], li[0], pj[0]) == orc.find_closest(lig, prot) == (1.0, 1, 0)


def test_reference_api_mirror_reads_like_the_go_tests(gpu, ddd):
    """metropolis_test.go, restated against the drop-in API (same names, same assertions)."""
    M = ddd.metropolis
    M.Seed(1)
    mk = lambda a: M.Molecule(xyzq=np.array(a, copy=True))
    protein = mk(mock_protein(1.0, -1.0))
    # TestSimulateMultipleLigands (:8-20)
    ligands = [mk(mock_ligand()) for _ in range(3)]
    minLigands, minEnergies = M.SimulateMultipleLigands(protein, ligands, 1000, True, 300.0, 4)
    assert len(minLigands) == len(ligands) == len(minEnergies)
    # TestSimulateMultipleLigandsParallel (:22-34)
    ligands = [mk(mock_ligand()) for _ in range(5)]
    minLigands, minEnergies = M.SimulateMultipleLigandsParallel(protein, ligands, 2000, True, 300.0, 2)
    assert len(minLigands) == 5 and len(minEnergies) == 5
    for l, e in zip(minLigands, minEnergies):
        assert M.CalculateEnergy(protein, l) == pytest.approx(e, rel=1e-12)
    # TestCalculateEnergyPositive / Negative (:36-56)
    assert M.CalculateEnergy(mk(mock_protein(1.0, 1.0)), mk(mock_ligand(1.0, 1.0))) > 0
    assert M.CalculateEnergy(mk(mock_protein(1.0, 1.0)), mk(mock_ligand(-1.0, -1.0))) < 0
    assert M.CalculateEnergy(mk(mock_protein(1.0, 1.0)), mk(mock_ligand(1.0, 1.0))) == pytest.approx(7361215932.1677294, rel=1e-15)
    # TestShiftLigandCloserByThreshold (:90-101)
    ligand = mk(mock_ligand())
    shifted = M.ShiftLigandCloserByThreshold(ligand, protein, 5.0)
    d1 = M.Distance(shifted.atoms[0].Position, protein.atoms[0].Position)
    d2 = M.Distance(shifted.atoms[1].Position, protein.atoms[1].Position)
    assert not (d1 > 5.0 and d2 > 5.0) and shifted is ligand
    # TestFindClosestAtomDistance (:140-148): exact
    distance, lp, pp = M.FindClosestAtomDistance(mk(mock_ligand()), mk(mock_protein(2.0, 4.0)))
    assert distance == math.sqrt(12.0) and (lp.X, pp.X) == (3.0, 1.0)
    # validateResults.go: CompareRMSD pipeline runs and CalculateRMSD of identical poses is 0
    lig = mk(mock_ligand())
    assert M.CalculateRMSD(lig, M.CopyLigand(lig)) == 0.0
    r = M.CompareRMSD(protein, mk(mock_ligand()), 400, False, M.TEMPERATURE, 4)
    assert np.isfinite(r) and r > 0
    # a single chain and the dead serial variant keep their signatures
    out = []
    M.SimulateEnergyMinimizationOneProc(protein, mk(mock_ligand()), 100, True, 300.0, out)
    assert len(out) == 1 and len(out[0]) == 2
    assert len(M.SimulateEnergyMinimization(protein, mk(mock_ligand()), 100, False, 300.0)) == 2
    # the caller's ligand is not mutated by the parallel entry points (each chain owns a private copy)
    keep = mk(mock_ligand())
    M.SimulateEnergyMinimizationParallel(protein, keep, 400, True, 300.0, 4)
    assert np.array_equal(keep.xyzq, mock_ligand())


def test_screen_rmsd_is_compare_rmsd_without_leaving_the_device(gpu, orc, synth_small):
    # CompareRMSD (validateResults.go:40-45): chains, then CalculateRMSD(final pose, start pose) per chain
    prot, off, lig = synth_small
    with gpu.Receptor(prot) as rec:
        scr = gpu.Screen(rec, off, lig, 3)
        scr.run(80, True, 310.15, gpu.default_params(), seed=4)
        poses, stats = scr.download()
        r, ms = scr.rmsd()
        scr.close()
    assert ms > 0 and r.shape == (3 * (len(off) - 1),)
    for c in range(len(r)):
        li, rep = divmod(c, 3)
        ref = lig[off[li]:off[li + 1]]
        want = orc.rmsd(gpu.chain_pose(off, 3, poses, li, rep), ref)
        assert r[c] == pytest.approx(want, rel=1e-13, abs=0) or (r[c] == 0.0 and want == 0.0)
    assert np.all((r > 0) == (stats["accepted"] > 0))
