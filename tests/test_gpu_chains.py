"""-m gpu: the Metropolis chain (SimulateEnergyMinimizationOneProc, metropolis.go:116-136) on device.

* F64 kernel driven by the same uniform stream as the oracle (Philox or a recorded array): identical accept /
  reject sequence, draws, retries, and final pose to 1e-9 A (sin/cos differ from libm by <= 1 ulp).
* F32 production kernel: same per-chain statistics as the oracle (acceptance counts, final-energy distribution).
* Replica pick (SimulateEnergyMinimizationParallel, :94-111) and the ligand batch (:27-67) vs oracle and goldens.
"""
import json

import numpy as np
import pytest

from conftest import GOLDEN, mock_ligand, mock_protein

pytestmark = pytest.mark.gpu
T = 310.15


def _p64(gpu, **kw):
    return gpu.default_params(precision=gpu.PRECISION_F64, **kw)


@pytest.mark.parametrize("rotate", [False, True])
@pytest.mark.parametrize("width", [32, 64, 128, 256])
def test_replay_matches_oracle_every_cta_width(gpu, orc, synth_small, rotate, width):
    prot, off, lig = synth_small
    with gpu.Receptor(prot) as rec:
        scr = gpu.Screen(rec, off, lig, 2)
        scr.set_chain_threads(width)
        scr.run(120, rotate, T, _p64(gpu), seed=9)
        poses, stats = scr.download()
        assert scr.chain_threads() == width
        scr.close()
    for c in range(2 * (len(off) - 1)):
        li, r = divmod(c, 2)
        pose_o, st_o = orc.chain(prot, lig[off[li]:off[li + 1]], 120, rotate, T, orc.philox_stream(9, c))
        for k in ("steps", "accepted", "downhill", "retries", "draws", "status"):
            assert stats[c][k] == st_o[k], (c, k)
        assert np.allclose(gpu.chain_pose(off, 2, poses, li, r), pose_o, rtol=0, atol=1e-9)
        assert stats[c]["final_energy"] == pytest.approx(st_o["final_energy"], rel=1e-11)
        assert stats[c]["start_energy"] == pytest.approx(st_o["start_energy"], rel=1e-12)


@pytest.mark.parametrize("rotate", [False, True])
@pytest.mark.parametrize("slots", [0, 1, 3, 16])
def test_replay_matches_oracle_sm_persistent_engine(gpu, orc, synth_small, rotate, slots):
    # the default engine: one 16-warp CTA per SM, `slots` resident chains, warps take pair-sum items of any chain
    prot, off, lig = synth_small
    with gpu.Receptor(prot) as rec:
        scr = gpu.Screen(rec, off, lig, 3)
        scr.set_chain_slots(slots)
        scr.run(120, rotate, T, _p64(gpu), seed=9)
        poses, stats = scr.download()
        assert scr.chain_threads() == 512 and scr.chain_slots() == (slots or 1)      # 24 chains on >= 24 SMs: 1 per SM
        scr.close()
    for c in range(3 * (len(off) - 1)):
        li, r = divmod(c, 3)
        pose_o, st_o = orc.chain(prot, lig[off[li]:off[li + 1]], 120, rotate, T, orc.philox_stream(9, c))
        for k in ("steps", "accepted", "downhill", "retries", "draws", "status"):
            assert stats[c][k] == st_o[k], (c, k)
        assert np.allclose(gpu.chain_pose(off, 3, poses, li, r), pose_o, rtol=0, atol=1e-9)
        assert stats[c]["final_energy"] == pytest.approx(st_o["final_energy"], rel=1e-11)
        assert stats[c]["start_energy"] == pytest.approx(st_o["start_energy"], rel=1e-12)


def test_engines_agree_and_queue_refills(gpu, ddd):
    # more chains than grid * slots: slots refill from the launch queue; F64 results are engine-independent bit for bit
    # except for the energy's summation order (rel 1e-12); mixed ligand sizes exercise the longest-first order
    prot = ddd.synth.make_receptor(700, seed=5)
    sizes = [3 + (7 * i) % 30 for i in range(400)]
    off, lig = ddd.synth.make_ligands(sizes, prot)
    with gpu.Receptor(prot) as rec:
        scr = gpu.Screen(rec, off, lig, 2)
        scr.set_chain_threads(64)
        scr.run(40, True, T, _p64(gpu), seed=12)
        p_cta, s_cta = scr.download()
        scr.set_chain_slots(2)
        scr.run(40, True, T, _p64(gpu), seed=12)
        p_sm, s_sm = scr.download()
        assert scr.chain_slots() == 2
        scr.set_chain_slots(0)
        scr.run(40, True, T, gpu.default_params(), seed=12)
        p32, s32 = scr.download()
        scr.run(40, True, T, gpu.default_params(), seed=12)
        p32b, s32b = scr.download()
        scr.close()
    for k in ("steps", "accepted", "downhill", "retries", "draws", "status"):
        assert np.array_equal(s_cta[k], s_sm[k]), k
    assert np.array_equal(p_cta, p_sm)
    assert np.allclose(s_cta["final_energy"], s_sm["final_energy"], rtol=1e-11)
    assert np.array_equal(p32, p32b) and np.array_equal(s32, s32b)                   # fp32 engine is bit-reproducible
    assert abs(s32["accepted"].sum() - s_sm["accepted"].sum()) <= 0.03 * s_sm["accepted"].sum()


@pytest.mark.parametrize("seg,slots,rotate", [(16, 0, True), (32, 1, False), (64, 3, True), (16, 16, True)])
def test_chain_segments_replay_the_oracle(gpu, orc, synth_small, seg, slots, rotate):
    # chain segments (ddd_screen_set_segment_steps): every `seg` steps a chain is parked in its output records and resumed by
    # whichever slot draws its queue position -- the parked state is exact, so the fp64 replay still matches the oracle step
    # for step.  Forced on a launch with as many slots as chains (or more): the spare slots wait on the ring.
    prot, off, lig = synth_small
    with gpu.Receptor(prot) as rec:
        scr = gpu.Screen(rec, off, lig, 3)
        scr.set_chain_slots(slots)
        scr.set_segment_steps(seg)
        scr.run(150, rotate, T, _p64(gpu), seed=9)
        poses, stats = scr.download()
        assert scr.segment_steps() == seg
        scr.close()
    for c in range(3 * (len(off) - 1)):
        li, r = divmod(c, 3)
        pose_o, st_o = orc.chain(prot, lig[off[li]:off[li + 1]], 150, rotate, T, orc.philox_stream(9, c))
        for k in ("steps", "accepted", "downhill", "retries", "draws", "status"):
            assert stats[c][k] == st_o[k], (c, k)
        assert np.allclose(gpu.chain_pose(off, 3, poses, li, r), pose_o, rtol=0, atol=1e-9)
        assert stats[c]["final_energy"] == pytest.approx(st_o["final_energy"], rel=1e-11)
        assert stats[c]["start_energy"] == pytest.approx(st_o["start_energy"], rel=1e-12)


def test_chain_segments_do_not_change_a_bit(gpu, ddd):
    # more chains than slots (the case segments exist for), mixed ligand sizes, production precision, best-ever tracker and
    # fused RMSD on: whole chains == segments of 8 / 16 steps == the automatic choice, bit for bit; chains that stop at the
    # retry bound void their remaining queue positions (nothing waits for them)
    prot = ddd.synth.make_receptor(700, seed=5)
    sizes = [3 + (7 * i) % 30 for i in range(500)]
    off, lig = ddd.synth.make_ligands(sizes, prot)
    bad = [3, 77, 401]                                        # impossible ligands: two atoms 0.01 A apart (metropolis.go:155 spins)
    for b in bad:
        lig[off[b] + 1, :3] = lig[off[b], :3] + np.array([0.01, 0.0, 0.0])
    prm = gpu.default_params(max_retries=40)
    res = {}
    with gpu.Receptor(prot) as rec:
        scr = gpu.Screen(rec, off, lig, 2)
        scr.set_chain_slots(2)
        scr.set_track_best(True)
        for seg in (0, 8, 16, -1):
            scr.set_segment_steps(seg)
            scr.run(70 if seg >= 0 else 2100, True, T, prm, seed=12)
            res[seg] = (scr.download(), scr.download_best(), scr.fused_rmsd(), scr.segment_steps())
        scr.set_segment_steps(0)
        scr.run(2100, True, T, prm, seed=12)
        whole_long = (scr.download(), scr.download_best(), scr.fused_rmsd(), scr.segment_steps())
        scr.close()
    assert res[0][3] == 0 and res[8][3] == 8 and res[16][3] == 16 and res[-1][3] == 128 and whole_long[3] == 0
    for seg, ref in ((8, res[0]), (16, res[0]), (-1, whole_long)):
        (p, st), (bp, be), rm, _ = res[seg]
        (p0, st0), (bp0, be0), rm0, _ = ref
        assert np.array_equal(p, p0) and np.array_equal(st, st0), seg
        assert np.array_equal(bp, bp0) and np.array_equal(be, be0) and np.array_equal(rm, rm0, equal_nan=True), seg
    st0 = res[0][0][1]
    stopped = np.nonzero(st0["status"] & gpu.STATUS_RETRY_LIMIT)[0]
    assert len(stopped) >= 3 and set(stopped // 2) <= set(bad) and np.all(st0["steps"][stopped] < 70)


def test_replay_golden_chains_on_shipped_data(gpu, sample):
    g = json.loads((GOLDEN / "oracle_golden.json").read_text())["chains"]
    with gpu.Receptor(sample["223l_protein"]) as rec:
        for c in g:
            lig = sample[c["ligand"] + "_ligand"]
            off = np.array([0, len(lig)])
            poses, stats, trace = gpu.run_chains(rec, off, lig, 1, c["steps"], c["rotate"], T, _p64(gpu), seed=c["seed"],
                                                 first_chain_id=c["chain_id"], want_trace=True)
            assert "".join(map(str, trace[0].tolist())) == c["accept_trace"]
            assert stats[0]["accepted"] == c["accepted"] and stats[0]["draws"] == c["draws"] and stats[0]["retries"] == c["retries"]
            assert stats[0]["final_energy"] == pytest.approx(c["final_energy"], rel=1e-11)
            assert np.allclose(poses[:, :3], np.array(c["pose"]), rtol=0, atol=1e-9)


def test_recorded_stream_replay(gpu, orc, synth_small):
    # the Go harness records rand.Float64() draws; the device must consume them in metropolis.go order
    prot, off, lig = synth_small
    sel = [0, 2, 4]
    ligs = [lig[off[i]:off[i + 1]] for i in sel]
    o2, l2 = gpu.pack_ligands(ligs)
    rng = np.random.default_rng(42)
    streams = [rng.random(60 * (4 + 3 * len(l) + 1) * 2) for l in ligs]
    roff = np.concatenate([[0], np.cumsum([len(s) for s in streams])]).astype(np.int64)
    with gpu.Receptor(prot) as rec:
        poses, stats, trace = gpu.run_chains(rec, o2, l2, 1, 60, True, T, _p64(gpu), seed=0, recorded=np.concatenate(streams),
                                             recorded_offsets=roff, want_trace=True)
        for c, l in enumerate(ligs):
            pose_o, st_o, tr_o, _ = orc.chain(prot, l, 60, True, T, orc.recorded_stream(streams[c]), want_trace=True)
            assert np.array_equal(trace[c], tr_o) and stats[c]["draws"] == st_o["draws"] and stats[c]["status"] == 0
            assert np.allclose(gpu.chain_pose(o2, 1, poses, c, 0), pose_o, rtol=0, atol=1e-9)
        # a stream that is too short is flagged, not silently replaced
        short = [s[:50] for s in streams]
        roff2 = np.concatenate([[0], np.cumsum([len(s) for s in short])]).astype(np.int64)
        _, stats, _ = gpu.run_chains(rec, o2, l2, 1, 60, True, T, _p64(gpu), seed=0, recorded=np.concatenate(short), recorded_offsets=roff2)
        assert np.all(stats["status"] & gpu.STATUS_STREAM_EXHAUSTED)


@pytest.mark.parametrize("temperature,k_coulomb", [(3.0e8, 9e9), (0.02, 1.0)])
@pytest.mark.parametrize("engine", ["sm", "cta64", "sm_slots3"])
def test_accept_branch_with_intermediate_probability(gpu, orc, synth_small, temperature, k_coulomb, engine):
    # AcceptMove's uphill branch (metropolis.go:145-148) with 0 < exp(-dE/T) < 1: at the reference constants (K = 9e9,
    # T = 310.15) the probability is always exactly 0 or 1, so here T is raised / K lowered until about half of
    # the uphill moves are accepted.  fp64 replay: every draw-vs-exp decision, the conditional draw count and the discarded
    # speculative proposals (an accept invalidates the speculation) must reproduce the oracle step for step.
    prot, off, lig = synth_small
    n = len(off) - 1
    prm = _p64(gpu, k_coulomb=k_coulomb)
    with gpu.Receptor(prot) as rec:
        scr = gpu.Screen(rec, off, lig, 2)
        if engine == "cta64":
            scr.set_chain_threads(64)
        elif engine == "sm_slots3":
            scr.set_chain_slots(3)
        scr.run(200, True, temperature, prm, seed=9)
        poses, stats = scr.download()
        scr.close()
        _, _, trace = gpu.run_chains(rec, off, lig, 2, 200, True, temperature, prm, seed=9, want_trace=True)
    po = orc.params(k_coulomb=k_coulomb)
    uphill_acc = rejected = 0
    for c in range(2 * n):
        li, r = divmod(c, 2)
        pose_o, st_o, tr_o, _ = orc.chain(prot, lig[off[li]:off[li + 1]], 200, True, temperature, orc.philox_stream(9, c), po, want_trace=True)
        for k in ("steps", "accepted", "downhill", "retries", "draws", "status"):
            assert stats[c][k] == st_o[k], (c, k)
        assert np.array_equal(trace[c], tr_o), c
        assert np.allclose(gpu.chain_pose(off, 2, poses, li, r), pose_o, rtol=0, atol=1e-9)
        assert stats[c]["final_energy"] == pytest.approx(st_o["final_energy"], rel=1e-10)
        uphill_acc += st_o["accepted"] - st_o["downhill"]
        rejected += st_o["steps"] - st_o["accepted"]
    assert uphill_acc > 200 and rejected > 200          # both outcomes of the draw < exp(x) comparison really occur


def test_segments_keep_recorded_streams_traces_and_uphill_accepts(gpu, orc, synth_small, monkeypatch):
    # DDD_SEGMENT_STEPS forces chain segments on the one-shot entry points too: a chain parked every 16 steps must resume at the
    # same position of a RECORDED stream, keep writing its accept trace, and take the same draw < exp(x) decisions (T = 3e8)
    monkeypatch.setenv("DDD_SEGMENT_STEPS", "16")
    prot, off, lig = synth_small
    sel = [0, 2, 4]
    ligs = [lig[off[i]:off[i + 1]] for i in sel]
    o2, l2 = gpu.pack_ligands(ligs)
    rng = np.random.default_rng(42)
    streams = [rng.random(90 * (4 + 3 * len(l) + 1) * 2) for l in ligs]
    roff = np.concatenate([[0], np.cumsum([len(s) for s in streams])]).astype(np.int64)
    with gpu.Receptor(prot) as rec:
        for temperature in (T, 3.0e8):
            poses, stats, trace = gpu.run_chains(rec, o2, l2, 1, 90, True, temperature, _p64(gpu), seed=0, recorded=np.concatenate(streams),
                                                 recorded_offsets=roff, want_trace=True)
            for c, l in enumerate(ligs):
                pose_o, st_o, tr_o, _ = orc.chain(prot, l, 90, True, temperature, orc.recorded_stream(streams[c]), want_trace=True)
                assert np.array_equal(trace[c], tr_o) and stats[c]["status"] == 0
                for k in ("steps", "accepted", "downhill", "retries", "draws"):
                    assert stats[c][k] == st_o[k], (c, k)
                assert np.allclose(gpu.chain_pose(o2, 1, poses, c, 0), pose_o, rtol=0, atol=1e-9)
        # and the replica pick on top of segmented chains (minimize_batch) still equals the unsegmented call
        a = gpu.minimize_batch(rec, off, lig, 240, True, T, 3, gpu.default_params(), seed=21)
        monkeypatch.setenv("DDD_SEGMENT_STEPS", "0")
        b = gpu.minimize_batch(rec, off, lig, 240, True, T, 3, gpu.default_params(), seed=21)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))


def test_accept_branch_intermediate_probability_recorded_stream_and_f32(gpu, orc, synth_small):
    prot, off, lig = synth_small
    sel = [2, 4, 7]
    ligs = [lig[off[i]:off[i + 1]] for i in sel]
    o2, l2 = gpu.pack_ligands(ligs)
    rng = np.random.default_rng(7)
    streams = [rng.random(80 * (4 + 3 * len(l) + 1) * 2) for l in ligs]
    roff = np.concatenate([[0], np.cumsum([len(s) for s in streams])]).astype(np.int64)
    Tt = 3.0e8
    with gpu.Receptor(prot) as rec:
        poses, stats, trace = gpu.run_chains(rec, o2, l2, 1, 80, True, Tt, _p64(gpu), seed=0, recorded=np.concatenate(streams),
                                             recorded_offsets=roff, want_trace=True)
        p32, s32, t32 = gpu.run_chains(rec, o2, l2, 1, 80, True, Tt, gpu.default_params(), seed=5, want_trace=True)
    for c, l in enumerate(ligs):
        pose_o, st_o, tr_o, _ = orc.chain(prot, l, 80, True, Tt, orc.recorded_stream(streams[c]), want_trace=True)
        assert np.array_equal(trace[c], tr_o) and stats[c]["draws"] == st_o["draws"] and stats[c]["status"] == 0
        assert st_o["accepted"] > st_o["downhill"]
        assert np.allclose(gpu.chain_pose(o2, 1, poses, c, 0), pose_o, rtol=0, atol=1e-9)
        # production precision in the same regime: same decisions unless an fp32 near-tie intervenes; statistics alike
        pose_p, st_p, tr_p, _ = orc.chain(prot, l, 80, True, Tt, orc.philox_stream(5, c), want_trace=True)
        assert abs(int(s32[c]["accepted"]) - int(st_p["accepted"])) <= 8
        assert s32[c]["accepted"] > s32[c]["downhill"]


def test_collision_retry_is_cumulative_and_bounded(gpu, orc):
    prot = np.array([[5.0, 5, 5, 1.0], [-3.0, 2, 1, -0.5]])
    lig = np.array([[0.0, 0, 0, 0.1], [0.52, 0, 0, -0.1], [0.0, 0.53, 0, 0.2]])      # pairs hover around MINDISTANCE
    off = np.array([0, 3])
    with gpu.Receptor(prot) as rec:
        poses, stats, trace = gpu.run_chains(rec, off, lig, 8, 200, False, T, _p64(gpu), seed=4, want_trace=True)
        tot = 0
        for c in range(8):
            pose_o, st_o, tr_o, _ = orc.chain(prot, lig, 200, False, T, orc.philox_stream(4, c), want_trace=True)
            assert np.array_equal(trace[c], tr_o) and stats[c]["retries"] == st_o["retries"] and stats[c]["draws"] == st_o["draws"]
            assert np.allclose(gpu.chain_pose(off, 8, poses, 0, c), pose_o, atol=1e-10)
            tot += st_o["retries"]
        assert tot > 0
        # impossible ligand: the reference would spin forever (metropolis.go:155); we stop, flag and restore the pose
        bad = np.array([[0.0, 0, 0, 0.1], [0.01, 0, 0, -0.1]])
        poses, stats, _ = gpu.run_chains(rec, np.array([0, 2]), bad, 1, 10, False, T, _p64(gpu, max_retries=5), seed=1)
        assert stats[0]["status"] & gpu.STATUS_RETRY_LIMIT and stats[0]["steps"] == 0 and stats[0]["retries"] == 5
        assert np.array_equal(poses, bad)


def test_zero_steps_and_tiny_ligands(gpu, orc, synth_small):
    prot, off, lig = synth_small
    with gpu.Receptor(prot) as rec:
        poses, stats, _ = gpu.run_chains(rec, off, lig, 3, 0, True, T, gpu.default_params(), seed=1)
        for c in range(3 * (len(off) - 1)):
            li, r = divmod(c, 3)
            assert np.array_equal(gpu.chain_pose(off, 3, poses, li, r), lig[off[li]:off[li + 1]])
            assert stats[c]["steps"] == 0 and stats[c]["final_energy"] == stats[c]["start_energy"]
        # an empty ligand in the batch: nothing to move, zero energy
        o2 = np.array([0, 0, 2]); l2 = mock_ligand()
        poses, stats, _ = gpu.run_chains(rec, o2, l2, 1, 20, True, T, _p64(gpu), seed=1)
        assert stats[0]["steps"] == 20 and stats[0]["final_energy"] == 0.0 and poses.shape == (2, 4)
        assert gpu.run_chains(rec, np.array([0]), np.zeros((0, 4)), 2, 10, True, T)[0].shape == (0, 4)


def test_empty_receptor_and_oversized_ligand(gpu, orc, ddd):
    # an empty protein: every energy is 0, nothing is ever downhill, the chain still consumes its draws (AcceptMove :141-149)
    lig = mock_ligand()
    with gpu.Receptor(np.zeros((0, 4))) as rec:
        for prm in (gpu.default_params(), _p64(gpu)):
            poses, stats, trace = gpu.run_chains(rec, np.array([0, 2]), lig, 2, 30, True, T, prm, seed=3, want_trace=True)
            pose_o, st_o, tr_o, _ = orc.chain(np.zeros((0, 4)), lig, 30, True, T, orc.philox_stream(3, 0), want_trace=True)
            assert np.array_equal(trace[0], tr_o) and stats[0]["draws"] == st_o["draws"] and stats[0]["final_energy"] == 0.0
            assert np.allclose(gpu.chain_pose(np.array([0, 2]), 2, poses, 0, 0), pose_o, atol=1e-9)
    # a ligand too large for a slot of the SM-persistent engine falls back to one CTA per chain, same results
    prot = ddd.synth.make_receptor(300, seed=1)
    rng = np.random.default_rng(0)
    big = np.concatenate([rng.uniform(-60, 60, size=(1800, 3)) * np.array([1.0, 1.0, 1.0]) + 200.0, rng.normal(0, 0.2, size=(1800, 1))], axis=1)
    with gpu.Receptor(prot) as rec:
        scr = gpu.Screen(rec, np.array([0, 1800]), big, 1)
        scr.run(3, True, T, _p64(gpu), seed=2)
        poses, stats = scr.download()
        assert scr.chain_slots() == 0 and stats[0]["steps"] == 3
        scr.close()
    pose_o, st_o = orc.chain(prot, big, 3, True, T, orc.philox_stream(2, 0))
    assert stats[0]["draws"] == st_o["draws"] and stats[0]["accepted"] == st_o["accepted"]
    assert np.allclose(poses, pose_o, rtol=0, atol=1e-8)


def test_receptor_sizes_cover_both_sum_layouts(gpu, ddd):
    # <= 8192 receptor atoms: per-group per-lane sums; above: one fp64 sum per static item.  Either way the chain's own
    # fp32-summed energy must agree with the fp64 energy of the same pose, and the run must be bit-reproducible.
    for n_rec in (130, 1288, 9000):
        prot = ddd.synth.make_receptor(n_rec, seed=8)
        off, lig = ddd.synth.make_ligands([40, 17, 33, 8] * 6, prot)
        with gpu.Receptor(prot) as rec:
            scr = gpu.Screen(rec, off, lig, 1)
            scr.run(60, True, T, gpu.default_params(), seed=5)
            poses, stats = scr.download()
            scr.run(60, True, T, gpu.default_params(), seed=5)
            poses2, stats2 = scr.download()
            scr.close()
            e64, ab = gpu.energy_batch(rec, off, poses, gpu.default_params(precision=gpu.PRECISION_F64), with_abs=True)
        assert np.array_equal(poses, poses2) and np.array_equal(stats, stats2), n_rec
        assert np.all(stats["status"] == 0) and np.all(stats["final_energy"] <= stats["start_energy"])
        assert np.all(np.abs(stats["final_energy"] - e64) <= 1e-13 * ab + 1e-300), n_rec
        assert np.all(np.abs(stats["chain_energy"] - e64) <= 1e-6 * ab + 1e-300), n_rec


def test_f32_chains_match_oracle_statistics(gpu, orc, ddd):
    # production precision: trajectories may differ after an fp32 near-tie, statistics must not
    prot = ddd.synth.make_receptor(1500, seed=2)
    off, lig = ddd.synth.make_ligands([24] * 48, prot)
    for rotate in (False, True):
        with gpu.Receptor(prot) as rec:
            poses, stats, trace = gpu.run_chains(rec, off, lig, 1, 250, rotate, T, gpu.default_params(), seed=17, want_trace=True)
        acc_o, same = [], []
        for c in range(48):
            pose_o, st_o, tr_o, _ = orc.chain(prot, lig[off[c]:off[c + 1]], 250, rotate, T, orc.philox_stream(17, c), want_trace=True)
            acc_o.append(st_o["accepted"])
            if np.array_equal(tr_o, trace[c]):                       # same decisions -> same trajectory
                same.append(c)
                assert np.allclose(gpu.chain_pose(off, 1, poses, c, 0), pose_o, rtol=0, atol=1e-9)
                assert stats[c]["final_energy"] == pytest.approx(st_o["final_energy"], rel=1e-9)
                assert stats[c]["draws"] == st_o["draws"] and stats[c]["retries"] == st_o["retries"]
        acc_o = np.array(acc_o)
        assert abs(stats["accepted"].mean() - acc_o.mean()) <= 0.02 * max(1.0, acc_o.mean())
        # measured on B200 (tools/f32_divergence.py, profiles/r02_f32_trace_divergence.json): 46 / 48 identical traces without
        # rotation (the two others part at steps 50-99 and 200-249: an fp32 near-tie), 48 / 48 with it
        assert len(same) >= 44
        # the chain's own fp32-summed energy agrees with the fp64 energy of the same pose (conditioning-aware)
        e, ab = gpu.energy_batch(gpu.Receptor(prot), off, poses, gpu.default_params(precision=gpu.PRECISION_F64), with_abs=True)
        assert np.all(np.abs(e - stats["final_energy"]) <= 1e-13 * ab)          # same fp64 terms, different summation order
        assert np.all(np.abs(stats["chain_energy"] - stats["final_energy"]) <= 1e-6 * ab)
        assert np.all(stats["status"] == 0)


@pytest.mark.parametrize("seed", [101, 102, 103])
def test_f32_statistics_match_across_seeds(gpu, orc, ddd, seed):
    # north_star: per-ligand minimum-energy and acceptance-rate statistics of the production (fp32) chains match the
    # float64 restatement across seeds; replica pick included (SimulateEnergyMinimizationParallel, :94-111)
    prot = ddd.synth.make_receptor(1200, seed=4)
    off, lig = ddd.synth.make_ligands([18, 25, 31, 40] * 6, prot)
    with gpu.Receptor(prot) as rec:
        out, en, picked, st = gpu.minimize_batch(rec, off, lig, 900, True, T, 3, gpu.default_params(), seed=seed, want_stats=True)
    rc, xo, eno, sto = orc.simulate_multiple_ligands(prot, off, lig, 900, True, T, 3, seed)
    assert rc == 0
    acc_g, acc_o = st["accepted"].astype(float), sto["accepted"].astype(float)
    assert abs(acc_g.mean() - acc_o.mean()) <= 0.02 * acc_o.mean() + 0.5                 # acceptance rate
    same = st["accepted"] == sto["accepted"]
    assert same.mean() >= 0.8                                                            # most chains never meet an fp32 near-tie
    assert np.allclose(st["final_energy"][same], sto["final_energy"][same], rtol=1e-9)
    # per-ligand minimised energies: identical where the picked chains agree, same distribution overall
    agree = np.isclose(en, eno, rtol=1e-9)
    assert agree.mean() >= 0.75
    assert abs(en.mean() - eno.mean()) <= 0.02 * abs(eno.mean())


def test_chain_energy_never_increases_at_this_temperature(gpu, ddd):
    # with K = 9e9 the sampler is greedy descent; accepted == downhill and E_final <= E_start
    prot = ddd.synth.make_receptor(1500, seed=2)
    off, lig = ddd.synth.make_ligands([24] * 64, prot)
    with gpu.Receptor(prot) as rec:
        _, stats, _ = gpu.run_chains(rec, off, lig, 2, 300, True, T, gpu.default_params(), seed=3)
    assert np.all(stats["final_energy"] <= stats["start_energy"])
    assert np.all(stats["accepted"] == stats["downhill"])
    assert np.all(stats["draws"] == 300 * (4 + 3 * 24) + (300 - stats["downhill"]) + stats["retries"] * (4 + 3 * 24))


def test_minimize_batch_matches_oracle(gpu, orc, synth_small):
    prot, off, lig = synth_small
    g = json.loads((GOLDEN / "oracle_golden.json").read_text())["synth600"]
    with gpu.Receptor(prot) as rec:
        for rotate, key in ((True, "minimize_rotate"), (False, "minimize_jitter")):
            out, en, picked, st = gpu.minimize_batch(rec, off, lig, 1000, rotate, T, 4, _p64(gpu), seed=21, want_stats=True)
            assert np.allclose(en, np.array(g[key]["energy"]), rtol=1e-10)
            assert st["accepted"].tolist() == g[key]["accepted"]
            assert np.all(st["steps"] == 250)                              # iterations / numProcs (:97)
            rc, xo, eno, sto = orc.simulate_multiple_ligands(prot, off, lig, 1000, rotate, T, 4, 21)
            assert np.allclose(out, xo, rtol=0, atol=1e-9)
            e64 = gpu.energy_batch(rec, off, out, _p64(gpu))
            assert np.allclose(en, e64, rtol=1e-12)                        # energy[i] = CalculateEnergy(final pose) (:61)
        # splitting the ligand list over two calls (or ranks) reproduces the single call
        a = gpu.minimize_batch(rec, off[:4], lig[:off[3]], 600, True, T, 3, _p64(gpu), seed=8)
        o_b = off[3:] - off[3]
        b = gpu.minimize_batch(rec, o_b, lig[off[3]:], 600, True, T, 3, _p64(gpu), seed=8, first_ligand_index=3)
        full = gpu.minimize_batch(rec, off, lig, 600, True, T, 3, _p64(gpu), seed=8)
        assert np.array_equal(np.concatenate([a[0], b[0]]), full[0]) and np.array_equal(np.concatenate([a[1], b[1]]), full[1])
        # numProcs > iterations: width 0, every chain returns the start pose (:97 truncating division)
        out, en, picked = gpu.minimize_batch(rec, off, lig, 3, True, T, 8, _p64(gpu), seed=1)
        assert np.array_equal(out, lig)
        with pytest.raises(gpu.DddError):
            gpu.minimize_batch(rec, off, lig, 100, True, T, 0)             # Go: integer divide by zero


def test_config1_shipped_data_golden(gpu, sample):
    # RunMultipleLigands (main.go:125-158): 223l protein x first 5 ligands, 3000 it, rotate, numProcs = 8
    g = json.loads((GOLDEN / "oracle_golden.json").read_text())["config1"]
    first5 = ["1a0i", "223l", "227l", "421p", "6tol"]
    off, lig = gpu.pack_ligands([sample[n + "_ligand"] for n in first5])
    with gpu.Receptor(sample["223l_protein"]) as rec:
        for seed in ("1", "2"):
            out, en, picked, st = gpu.minimize_batch(rec, off, lig, 3000, True, T, 8, _p64(gpu), seed=int(seed), want_stats=True)
            assert st["accepted"].tolist() == g[seed]["accepted"] and st["retries"].tolist() == g[seed]["retries"]
            assert st["draws"].tolist() == g[seed]["draws"]
            assert np.allclose(en, g[seed]["energy"], rtol=1e-10)
            cs = [float(out[off[i]:off[i + 1], :3].sum()) for i in range(5)]
            assert np.allclose(cs, g[seed]["pose_checksum"], rtol=0, atol=1e-7)
            # production precision on the same run: same statistics
            out32, en32, _, st32 = gpu.minimize_batch(rec, off, lig, 3000, True, T, 8, gpu.default_params(), seed=int(seed), want_stats=True)
            assert abs(st32["accepted"].sum() - st["accepted"].sum()) <= 0.05 * st["accepted"].sum() + 2


def test_rigid_extension_matches_its_oracle(gpu, orc, synth_small):
    # north_star's rigid-body move has no reference row; parity is against the builder's own restatement only
    prot, off, lig = synth_small
    with gpu.Receptor(prot) as rec:
        poses, stats, trace = gpu.run_chains(rec, off, lig, 1, 150, True, T, _p64(gpu, move_mode=gpu.MOVE_RIGID), seed=6, want_trace=True)
    po = orc.params(move_mode=1)
    for c in range(len(off) - 1):
        l = lig[off[c]:off[c + 1]]
        pose_o, st_o, tr_o, _ = orc.chain(prot, l, 150, True, T, orc.philox_stream(6, c), po, want_trace=True)
        assert np.array_equal(trace[c], tr_o) and stats[c]["draws"] == st_o["draws"]
        g = gpu.chain_pose(off, 1, poses, c, 0)
        assert np.allclose(g, pose_o, atol=1e-9)
        if len(l) > 1:                                                      # rigid: intra-ligand distances preserved
            d0 = np.linalg.norm(l[:, None, :3] - l[None, :, :3], axis=-1)
            d1 = np.linalg.norm(g[:, None, :3] - g[None, :, :3], axis=-1)
            assert np.allclose(d0, d1, atol=1e-9)


def test_full_size_config2_properties(gpu, ddd):
    # BASELINE config 2 shape (5k receptor x 1024 x 40), shortened chains: size-independent invariants
    prot = ddd.synth.make_receptor(5000)
    off, lig = ddd.synth.make_ligands([40] * 1024, prot)
    with gpu.Receptor(prot) as rec:
        scr = gpu.Screen(rec, off, lig, 1)
        scr.run(200, True, T, gpu.default_params(), seed=1)
        poses, stats = scr.download()
        ms = scr.last_kernel_ms()
        scr.run(200, True, T, gpu.default_params(), seed=1)                 # idempotent: same seed -> same bits
        poses2, stats2 = scr.download()
        scr.close()
        e64, ab = gpu.energy_batch(rec, off, poses, gpu.default_params(precision=gpu.PRECISION_F64), with_abs=True)
    assert ms > 0 and np.array_equal(poses, poses2) and np.array_equal(stats, stats2)
    assert np.all(stats["steps"] == 200) and np.all(stats["status"] == 0)
    assert np.all(stats["final_energy"] <= stats["start_energy"])
    assert np.all(np.abs(e64 - stats["final_energy"]) <= 1e-13 * ab)              # same fp64 terms, different summation order
    assert np.all(np.abs(stats["chain_energy"] - e64) <= 1e-6 * ab)
    assert np.array_equal(poses[:, 3], lig[:, 3])                           # charges untouched
    moved = np.abs(poses[:, :3] - lig[:, :3]).reshape(1024, -1).max(1) > 0
    assert np.array_equal(moved, stats["accepted"] > 0)
