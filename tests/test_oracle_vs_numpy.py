"""Cross-checks the C oracle against an independently written numpy / pure-Python restatement
(tests/_np_restatement.py) where the reference's tests pin nothing, and against SURVEY.md App. B."""
import math

import numpy as np
import pytest

import _np_restatement as npr

PHILOX_KATS = [   # Random123 kat_vectors, philox4x32-10
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
     [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


@pytest.mark.parametrize("ctr,key,want", PHILOX_KATS)
def test_philox_known_answers(orc, ctr, key, want):
    assert orc.philox(ctr, key) == want
    assert npr.philox4x32_10(ctr, key) == want


def test_uniform_stream_matches_and_is_in_unit_interval(orc):
    s = orc.philox_stream(12345, 77)
    u = [orc.next(s) for _ in range(64)]
    assert u == [npr.u01(12345, 77, i) for i in range(64)]
    assert all(0.0 <= x < 1.0 for x in u)
    assert orc.u01(12345, 77, 63) == u[63]                 # counter-based: random access == sequential


def test_energy_bit_exact_small(orc, synth_small):
    prot, off, lig = synth_small
    for i in range(len(off) - 1):
        l = lig[off[i]:off[i + 1]]
        assert orc.energy(prot[:150], l) == npr.energy(prot[:150], l)     # same order -> same bits


def test_energy_sample_data_matches_survey_appendix_b(orc, sample):
    want = {"1a0i": -3671536276.9907126, "223l": 148197363.11284322, "227l": 32255670.150091246,
            "421p": -2858858144.3740206, "6tol": 24148058.84069746, "721p": -2873092883.6844983,
            "7loc": 332588.37324483041, "821p": -2872379867.0384798}
    for name, e in want.items():
        got = orc.energy(sample["223l_protein"], sample[name + "_ligand"])
        assert got == pytest.approx(e, rel=1e-15)
        ev, ab = npr.energy_vec(sample["223l_protein"], sample[name + "_ligand"])
        assert abs(got - ev) <= 1e-12 * ab
    # conditioning of the neutral ligand
    e, ab = orc.energy(sample["223l_protein"], sample["7loc_ligand"], with_abs=True)
    assert 4e5 < ab / abs(e) < 7e5


def test_energy_clamp(orc):
    prot = np.array([[1.0, 2.0, 3.0, 0.5]])
    lig = np.array([[1.0, 2.0, 3.0, -0.25]])
    assert orc.energy(prot, lig) == 9e9 * (0.5 * -0.25) / 1e-6           # metropolis.go:255-257


def test_empty_inputs(orc):
    prot = np.zeros((0, 4)); lig = np.array([[0.0, 0, 0, 1]])
    assert orc.energy(prot, lig) == 0.0 and orc.energy(lig, prot) == 0.0
    assert orc.is_collision_free(prot, 0.5)
    assert math.isnan(orc.rmsd(prot, prot))                              # 0/0, as in Go
    assert math.isnan(orc.rmsd(np.ones((3, 4)), np.ones((2, 4))))        # Go: index out of range


@pytest.mark.parametrize("rotate", [False, True])
def test_chain_matches_numpy_restatement(orc, synth_small, rotate):
    prot, off, lig = synth_small
    prot = prot[:120]
    l = lig[off[0]:off[1]]
    pose, st, tr, _ = orc.chain(prot, l, 60, rotate, 310.15, orc.philox_stream(9, 4), want_trace=True)
    pose_n, cur_n, tr_n = npr.chain(prot, l, 60, rotate, 310.15, npr.Stream(9, 4))
    assert tr.tolist() == tr_n
    assert np.allclose(pose[:, :3], pose_n[:, :3], rtol=0, atol=1e-11)
    assert st["final_energy"] == pytest.approx(cur_n, rel=1e-12)
    assert st["accepted"] == sum(tr_n)


@pytest.mark.parametrize("rotate", [False, True])
def test_chain_uphill_accept_branch_matches_numpy_restatement(orc, synth_small, rotate):
    # AcceptMove :145-148 with 0 < exp(-dE/T) < 1 (T raised to 3e8: at the reference constants p is always 0 or 1)
    prot, off, lig = synth_small
    prot = prot[:120]
    l = lig[off[2]:off[3]]
    T = 3.0e8 if rotate else 1.0e7                                        # jitter-only moves change the energy far less
    pose, st, tr, _ = orc.chain(prot, l, 80, rotate, T, orc.philox_stream(9, 4), want_trace=True)
    pose_n, cur_n, tr_n = npr.chain(prot, l, 80, rotate, T, npr.Stream(9, 4))
    assert tr.tolist() == tr_n
    assert np.allclose(pose[:, :3], pose_n[:, :3], rtol=0, atol=1e-11)
    assert st["final_energy"] == pytest.approx(cur_n, rel=1e-12)
    assert st["accepted"] > st["downhill"] and st["accepted"] < 80        # uphill moves both accepted and rejected


def test_chain_recorded_stream_equals_philox_stream(orc, synth_small):
    prot, off, lig = synth_small
    l = lig[off[2]:off[3]]
    pose, st, tr, _ = orc.chain(prot, l, 40, True, 310.15, orc.philox_stream(5, 1), want_trace=True)
    u = np.array([orc.u01(5, 1, i) for i in range(int(st["draws"]))])
    pose2, st2, tr2, _ = orc.chain(prot, l, 40, True, 310.15, orc.recorded_stream(u), want_trace=True)
    assert np.array_equal(pose, pose2) and tr.tolist() == tr2.tolist() and st2["status"] == 0
    # one uniform short: the stream runs dry and is flagged
    *_, = orc.chain(prot, l, 40, True, 310.15, orc.recorded_stream(u[:-1]))
    assert orc.chain(prot, l, 40, True, 310.15, orc.recorded_stream(u[:-1]))[1]["status"] & 2


def test_collision_retry_is_cumulative(orc):
    # two atoms 0.52 A apart: some jitters bring them under 0.5 -> the retry re-jitters the moved atoms
    lig = np.array([[0.0, 0, 0, 0.1], [0.52, 0, 0, -0.1]])
    total_rej = 0
    for seed in range(40):
        s = orc.philox_stream(seed, 0)
        out, rej = orc.jitter_ligand(lig, s)
        r = npr.Stream(seed, 0)
        l2 = lig.copy()
        rej2 = npr.propose(l2, False, r)
        assert rej == rej2 and s.pos == r.pos == 6 * (rej + 1)
        assert np.allclose(out, l2, atol=1e-15)
        assert orc.is_collision_free(out, 0.5)
        total_rej += rej
    assert total_rej > 0


def test_retry_bound(orc):
    lig = np.array([[0.0, 0, 0, 0.1], [0.01, 0, 0, -0.1]])               # can never get 0.5 apart by +-0.05 jitters soon
    out, rej = orc.jitter_ligand(lig, orc.philox_stream(1, 0), orc.params(max_retries=5))
    assert rej == -6                                                       # -(5) - 1: bound hit
    pose, st = orc.chain(np.array([[5.0, 5, 5, 1]]), lig, 10, False, 300.0, orc.philox_stream(1, 0), orc.params(max_retries=5))
    assert st["status"] & 1 and st["steps"] == 0 and np.array_equal(pose, lig)


def test_minimize_parallel_matches_numpy(orc, synth_small):
    prot, off, lig = synth_small
    prot = prot[:100]
    l = lig[off[1]:off[2]]
    rc, out, st, poses, picked = orc.minimize_parallel(prot, l, 90, False, 310.15, 3, seed=2, lig_index=6)
    out_n, e_n = npr.minimize_parallel(prot, l, 90, False, 310.15, 3, 2, 6)
    assert rc == 0 and np.allclose(out[:, :3], out_n[:, :3], atol=1e-12)
    assert orc.energy(prot, out) == pytest.approx(e_n, rel=1e-13)
    assert np.all(st["steps"] == 30)


def test_rmsd_and_closest_match_numpy(orc, synth_small):
    prot, off, lig = synth_small
    for i in range(len(off) - 1):
        l = lig[off[i]:off[i + 1]]
        r = orc.randomize_pose(l, orc.philox_stream(3, i))
        assert orc.rmsd(r, l) == npr.rmsd(r, l)
        assert orc.find_closest(l, prot) == npr.find_closest(l, prot)


def test_randomize_pose_draw_order(orc, synth_small):
    prot, off, lig = synth_small
    l = lig[off[0]:off[1]].copy()
    r = npr.Stream(8, 2)
    want = l.copy()
    for a in want:                                                         # validateResults.go:71-76
        a[0] += (r() - 0.5) * 10.0; a[1] += (r() - 0.5) * 10.0; a[2] += (r() - 0.5) * 10.0
    npr.rotate(want, math.pi, r)                                           # :78
    got = orc.randomize_pose(l, orc.philox_stream(8, 2))
    assert np.allclose(got, want, atol=1e-12) and r.pos == 3 * len(l) + 4


def test_rigid_move_preserves_shape(orc, synth_small):
    prot, off, lig = synth_small
    l = lig[off[4]:off[5]]
    out = orc.rigid_move(l, orc.philox_stream(1, 0), orc.params(move_mode=1))
    d0 = np.linalg.norm(l[:, None, :3] - l[None, :, :3], axis=-1)
    d1 = np.linalg.norm(out[:, None, :3] - out[None, :, :3], axis=-1)
    assert np.allclose(d0, d1, atol=1e-12)
    assert np.linalg.norm(out[:, :3].mean(0) - l[:, :3].mean(0)) <= 0.5 * math.sqrt(3) / 2 + 1e-12


def test_pow_timing_variant_agrees_to_rounding(orc, synth_small):
    # the pow() build exists to mimic the COST of Go's out-of-line math.Pow in the CPU baseline; glibc pow is
    # not always correctly rounded, so it is not the oracle -- it must only agree to a few ulp of sum|term|
    from _oracle import Oracle
    powv = Oracle("pow")
    prot, off, lig = synth_small
    assert "pow" in powv.build_info() and "pow" not in orc.build_info()
    for i in range(3):
        l = lig[off[i]:off[i + 1]]
        e, ab = orc.energy(prot, l, with_abs=True)
        assert abs(powv.energy(prot, l) - e) <= 1e-15 * ab
