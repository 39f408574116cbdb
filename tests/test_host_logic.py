"""Host-side logic that needs no GPU: sharding plan, CSR packing, the scalar helpers of the
reference API mirror (metropolis.py) against the oracle, and the mol2 parser."""
import math

import numpy as np
import pytest

from conftest import mock_ligand, mock_protein


def test_shard_plan_uniform(capi):
    off = np.arange(0, 40 * 1025, 40)
    b = capi.shard_plan(off, 8, 8)
    assert b[0] == 0 and b[-1] == 1024 * 8 and np.all(np.diff(b) == 1024)
    assert capi.shard_plan(off, 1, 1).tolist() == [0, 1024]
    assert capi.shard_plan(off, 1, 3).tolist() == [0, 341, 683, 1024]


def test_shard_plan_balances_atoms_not_counts(capi, ddd):
    sizes = ddd.synth.sweep_sizes(5000)
    off = np.concatenate([[0], np.cumsum(sizes)])
    b = capi.shard_plan(off, 1, 8)
    assert np.all(np.diff(b) > 0) and b[-1] == 5000
    cost = np.array([off[b[k + 1]] - off[b[k]] + (b[k + 1] - b[k]) for k in range(8)], dtype=float)
    assert cost.max() / cost.mean() < 1.02
    counts = np.diff(b)
    assert counts.max() != counts.min()                     # equal-count sharding would be unbalanced


def test_shard_plan_edge_cases(capi):
    assert capi.shard_plan(np.array([0]), 4, 3).tolist() == [0, 0, 0, 0]            # no ligands
    b = capi.shard_plan(np.array([0, 10]), 8, 3)                                    # one ligand, 8 replicas: they stay together
    assert b[0] == 0 and b[-1] == 8 and sorted(np.diff(b).tolist()) == [0, 0, 8]
    b = capi.shard_plan(np.array([0, 10, 20, 30, 40, 50]), 8, 2)                    # cuts fall between ligands (replica pick per device)
    assert b.tolist() in ([0, 16, 40], [0, 24, 40])
    b = capi.shard_plan(np.array([0, 5, 5, 9]), 2, 8)                               # more shards than chains
    assert b[0] == 0 and b[-1] == 6 and np.all(np.diff(b) >= 0)
    with pytest.raises(capi.DddError):
        capi.shard_plan(np.array([0, 3]), 0, 2)
    with pytest.raises(capi.DddError):
        capi.shard_plan(np.array([0, 3, 2]), 1, 2)                                  # not monotone


def test_bench_rank_blocks_cover_the_workload_and_balance_config5(ddd):
    # bench.py's per-rank ligand blocks: contiguous, disjoint, complete for every config and world size; config 5 (20..80-atom
    # ligands) follows the library's own cost-balanced shard plan instead of an equal count
    import importlib.util
    from pathlib import Path
    spec = importlib.util.spec_from_file_location("ddd_bench", Path(__file__).resolve().parent.parent / "bench.py")
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for config, total in ((3, 4096), (5, 100000), (5, 1003)):
        for world in (1, 2, 4, 8):
            blocks = [bench.workload(config, r, world, total if config == 5 else 0)[:2] for r in range(world)]
            assert blocks[0][0] == 0 and sum(n for _, n in blocks) == total
            assert all(blocks[r][0] + blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
    sizes = ddd.synth.sweep_sizes(100000)
    work = [int(sizes[lo:lo + n].sum() + n) for lo, n in (bench.workload(5, r, 8, 0)[:2] for r in range(8))]
    assert max(work) <= 1.001 * np.mean(work)                                     # an equal count is 0.8 % off
    assert bench.workload(2, 3, 8, 0)[:3] == (3 * 1024, 1024, 1)                   # config 2 / 4: weak scaling, 1024 ligands per rank


def test_pack_unpack_roundtrip(capi):
    ligs = [np.random.default_rng(i).normal(size=(n, 4)) for i, n in enumerate([3, 0, 7, 1])]
    off, x = capi.pack_ligands(ligs)
    assert off.tolist() == [0, 3, 3, 10, 11] and x.shape == (11, 4)
    for a, b in zip(ligs, capi.unpack_ligands(off, x)):
        assert np.array_equal(a, b)
    off, x = capi.pack_ligands([])
    assert off.tolist() == [0] and x.shape == (0, 4)


def test_chain_pose_indexing(capi):
    off = np.array([0, 2, 5])
    poses = np.arange((2 + 3) * 3 * 4, dtype=float).reshape(-1, 4)          # replicas = 3, chain-major
    assert capi.chain_pose(off, 3, poses, 0, 2)[0, 0] == 4 * 4               # atoms 4..5
    assert capi.chain_pose(off, 3, poses, 1, 0)[0, 0] == 6 * 4 and len(capi.chain_pose(off, 3, poses, 1, 1)) == 3


# ---- reference API mirror: scalar helpers vs the oracle, fed the same recorded stream
def _mol(ddd, a):
    return ddd.metropolis.Molecule(xyzq=np.array(a, copy=True))


def test_mirror_constants(ddd):
    M = ddd.metropolis
    assert (M.K, M.THRESHOLD, M.MINDISTANCE, M.MAXANGLE, M.TEMPERATURE) == (9e9, 5, 0.5, 0.79, 310.15)


def test_mirror_distance_and_rotate_atom(ddd, orc):
    M = ddd.metropolis
    assert M.Distance(M.Position3d(1, 2, 3), M.Position3d(4, 6, 8)) == math.sqrt(50)          # TestDistance
    r = M.RotateAtom(M.Position3d(1, 0, 0), M.Position3d(0, 0, 1), math.pi / 2)                # TestRotateAtom
    assert abs(r.X) <= 1e-6 and abs(r.Y - 1) <= 1e-6 and abs(r.Z) <= 1e-6
    o = orc.rotate_atom([1.5, -2, 0.25], orc.normalize([1, 2, 3]), 0.3)
    ax = M.Position3d(1, 2, 3); ax.Normalize()
    r = M.RotateAtom(M.Position3d(1.5, -2, 0.25), ax, 0.3)
    assert [r.X, r.Y, r.Z] == o.tolist()


def test_mirror_normalize_zero_is_noop(ddd):
    v = ddd.metropolis.Position3d(0, 0, 0)
    v.Normalize()
    assert (v.X, v.Y, v.Z) == (0, 0, 0)                                                        # datatypes.go:35


def test_mirror_moves_match_oracle_on_recorded_stream(ddd, orc, synth_small):
    M = ddd.metropolis
    prot, off, lig = synth_small
    l = lig[off[2]:off[3]]
    u = np.array([orc.u01(3, 9, i) for i in range(400)])
    it = iter(u)
    m = M.JitterAndRotateLigand(_mol(ddd, l), M.MINDISTANCE, M.MAXANGLE, rand=lambda: next(it))
    want, rej = orc.jitter_and_rotate_ligand(l, orc.recorded_stream(u))
    assert rej == 0 and np.allclose(m.xyzq, want, atol=1e-13)
    it = iter(u)
    m = M.JitterLigand(_mol(ddd, l), M.MINDISTANCE, rand=lambda: next(it))
    assert np.allclose(m.xyzq, orc.jitter_ligand(l, orc.recorded_stream(u))[0], atol=1e-15)
    it = iter(u)
    m = M.RandomizeLigandPose(_mol(ddd, l), rand=lambda: next(it))
    assert np.allclose(m.xyzq, orc.randomize_pose(l, orc.recorded_stream(u)), atol=1e-12)


def test_mirror_in_place_aliasing(ddd):
    M = ddd.metropolis
    lig = _mol(ddd, mock_ligand())
    before = lig.xyzq.copy()
    out = M.RotateLigand(lig, math.pi / 4)
    assert out is lig and not np.array_equal(lig.xyzq, before)      # the caller's atoms are mutated
    cp = M.CopyLigand(lig)
    assert cp.xyzq is not lig.xyzq and np.array_equal(cp.xyzq, lig.xyzq)     # TestCopyLigand


def test_mirror_accept_move_and_collision(ddd):
    M = ddd.metropolis
    assert M.AcceptMove(10.0, 5.0, 300.0)                                    # TestAcceptMove
    assert M.AcceptMove(5.0, 5.0, 300.0, rand=lambda: 0.999)
    assert not M.AcceptMove(5.0, 305.0, 300.0, rand=lambda: 0.5)
    assert M.IsCollisionFree(_mol(ddd, mock_ligand()), 0.5)
    assert not M.IsCollisionFree(_mol(ddd, [[0, 0, 0, 1], [0.3, 0, 0, 1]]), 0.5)
    j = M.JitterLigand(_mol(ddd, mock_ligand()), 0.5)
    assert M.IsCollisionFree(j, 0.5)                                          # TestJitterLigand


def test_mirror_errors(ddd):
    M = ddd.metropolis
    p, l = _mol(ddd, mock_protein(1, -1)), _mol(ddd, mock_ligand())
    with pytest.raises(ZeroDivisionError):
        M.SimulateEnergyMinimizationParallel(p, l, 100, True, 300.0, 0)
    with pytest.raises(ZeroDivisionError):
        M.SimulateMultipleLigandsParallel(p, [l], 100, True, 300.0, 0)
    with pytest.raises(IndexError):
        M.CalculateRMSD(_mol(ddd, np.zeros((3, 4))), _mol(ddd, np.zeros((2, 4))))
    with pytest.raises(OSError):
        M.ParseMol2("/nonexistent/file.mol2")
    assert M.average([]) == 0 and M.average([1.0, 2.0]) == 1.5


MOL2 = """@<TRIPOS>MOLECULE
x.pdb
 3 2 0 0 0
SMALL
GASTEIGER

@<TRIPOS>ATOM
      1 CL         31.4870   20.3560   -8.2500 Cl    178  CL178       0.0000
      2  C1        43.9850   16.8300   32.3760 C.3   900  BME900      0.1946
      3  O1        45.0880   17.1670   31.5440 O.3   900  BME900     -0.2194
   short line with too few fields
      4  X1        bad       1.0       2.0    X     1    R1   oops
@<TRIPOS>BOND
     1    1    2   1
      9  C9         1.0  2.0  3.0 C.3   900  BME900      0.5
"""


def test_parse_mol2_mirror_and_oracle(ddd, orc, tmp_path):
    f = tmp_path / "t_ligand.mol2"
    f.write_text(MOL2)
    want = np.array([[31.487, 20.356, -8.25, 0.0], [43.985, 16.83, 32.376, 0.1946], [45.088, 17.167, 31.544, -0.2194],
                     [0.0, 1.0, 2.0, 0.0]])                  # unparsable numbers become 0 (parsing.go:85-88)
    assert np.array_equal(ddd.metropolis.ParseMol2(f).xyzq, want)
    assert np.array_equal(orc.parse_mol2(f), want)
    assert ddd.metropolis.ExtractFileLabel(str(f)) == "t"
    (tmp_path / "sub").mkdir(); (tmp_path / "sub" / "a_ligand.mol2").write_text(MOL2); (tmp_path / "b_protein.mol2").write_text(MOL2)
    found = ddd.metropolis.findFilesWithSubstring(str(tmp_path), "ligand")
    assert [p.split("/")[-1] for p in found] == ["a_ligand.mol2", "t_ligand.mol2"]


def test_synthetic_generators_are_deterministic_and_to_spec(ddd):
    s = ddd.synth
    r1, r2 = s.make_receptor(800), s.make_receptor(800)
    assert np.array_equal(r1, r2) and r1.shape == (800, 4)
    assert np.all(np.abs(r1[:, 3]) <= 0.55) and np.array_equal(r1, np.round(r1, 4))
    from scipy.spatial import cKDTree
    d, _ = cKDTree(r1[:, :3]).query(r1[:, :3], k=2)
    assert d[:, 1].min() >= 1.19
    off, lig = s.make_ligands([40, 21], r1)
    assert off.tolist() == [0, 40, 61]
    assert abs(lig[:40, 3].sum()) < 2e-3 and abs(lig[40:, 3].sum() + 1.0) < 2e-3      # net 0 (even) / -1 (odd)
    dd, _ = cKDTree(r1[:, :3]).query(lig[:, :3], k=1)
    assert 4.0 < dd[:40].min() <= 5.01
    sizes = s.sweep_sizes(1000)
    assert sizes.min() >= 20 and sizes.max() <= 80
