"""Cutoff extension (params.cutoff > 0): shifted-Coulomb truncation + uniform-grid cell list (BASELINE config 4).

The reference sums every pair (metropolis.go:249-259) and has no cutoff: this mode has NO reference row, its parity is
against the builder's own CPU restatement only ("parity unpinned"), and it is off by default -- the default path is
checked to be untouched.  E_cut = sum_{d < rc} K q_p q_l (1/max(d, clamp) - 1/rc).
"""
import numpy as np
import pytest

T = 310.15
K = 9e9


def _np_cut_energy(prot, lig, rc, clamp=1e-6):
    d = np.linalg.norm(prot[:, None, :3] - lig[None, :, :3], axis=-1)
    qq = prot[:, None, 3] * lig[None, :, 3]
    m = d < rc
    dc = np.maximum(d, clamp)
    return float((K * qq / dc - K * qq / rc)[m].sum()), float(np.abs(K * qq / dc - K * qq / rc)[m].sum())


# ------------------------------------------------------------------ CPU: the restatement itself
def test_oracle_cutoff_matches_numpy_and_defaults_off(orc, ddd):
    prot = ddd.synth.make_receptor(900, seed=3)
    off, lig = ddd.synth.make_ligands([17, 30], prot)
    for i in range(2):
        l = lig[off[i]:off[i + 1]]
        assert orc.params().cutoff == 0.0
        plain = orc.energy(prot, l)
        for rc in (6.0, 12.0, 25.0):
            e, ab = orc.energy(prot, l, orc.params(cutoff=rc), with_abs=True)
            en, abn = _np_cut_energy(prot, l, rc)
            assert e == pytest.approx(en, rel=0, abs=1e-12 * abn) and ab == pytest.approx(abn, rel=1e-12)
        # a cutoff beyond every pair: the full sum minus the constant shift K sum(q_p q_l) / rc
        rc = 1e4
        e = orc.energy(prot, l, orc.params(cutoff=rc))
        assert e == pytest.approx(plain - K * prot[:, 3].sum() * l[:, 3].sum() / rc, rel=1e-9)
    far = lig[off[0]:off[1]] + np.array([500.0, 0, 0, 0])
    assert orc.energy(prot, far, orc.params(cutoff=12.0)) == 0.0                 # nothing in reach
    # a chain under the cutoff is still greedy descent and consumes the same draws per attempt
    pose, st = orc.chain(prot, lig[off[0]:off[1]], 150, True, T, orc.philox_stream(5, 0), orc.params(cutoff=12.0))
    assert st["final_energy"] <= st["start_energy"] and st["steps"] == 150


def test_oracle_cutoff_golden(orc, sample):
    import json
    from conftest import GOLDEN
    g = json.loads((GOLDEN / "oracle_golden.json").read_text())["cutoff10"]
    pc = orc.params(cutoff=g["cutoff"])
    for l, v in g["energy"].items():
        e, a = orc.energy(sample["223l_protein"], sample[l + "_ligand"], pc, with_abs=True)
        assert e == v["energy"] and a == v["abs_sum"]
    for c in g["chains"]:
        pose, st, tr, _ = orc.chain(sample["223l_protein"], sample[c["ligand"] + "_ligand"], c["steps"], c["rotate"], T,
                                    orc.philox_stream(c["seed"], c["chain_id"]), pc, want_trace=True)
        assert "".join(map(str, tr.tolist())) == c["accept_trace"] and st["draws"] == c["draws"] and st["final_energy"] == c["final_energy"]


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_cutoff_golden_on_shipped_data(gpu, sample):
    # the committed restatement outputs, reproduced on the GPU box without the CPU oracle in the loop
    import json
    from conftest import GOLDEN
    g = json.loads((GOLDEN / "oracle_golden.json").read_text())["cutoff10"]
    p64 = gpu.default_params(precision=gpu.PRECISION_F64, cutoff=g["cutoff"])
    with gpu.Receptor(sample["223l_protein"]) as rec:
        for l, v in g["energy"].items():
            lig = sample[l + "_ligand"]
            e, a = gpu.energy_batch(rec, np.array([0, len(lig)]), lig, p64, with_abs=True)
            assert abs(e[0] - v["energy"]) <= 1e-13 * v["abs_sum"] and a[0] == pytest.approx(v["abs_sum"], rel=1e-12)
        for c in g["chains"]:
            lig = sample[c["ligand"] + "_ligand"]
            poses, stats, trace = gpu.run_chains(rec, np.array([0, len(lig)]), lig, 1, c["steps"], c["rotate"], T, p64, seed=c["seed"],
                                                 first_chain_id=c["chain_id"], want_trace=True)
            assert "".join(map(str, trace[0].tolist())) == c["accept_trace"]
            assert stats[0]["draws"] == c["draws"] and stats[0]["retries"] == c["retries"] and stats[0]["accepted"] == c["accepted"]
            assert stats[0]["final_energy"] == pytest.approx(c["final_energy"], rel=1e-10)
            assert float(poses[:, :3].sum()) == pytest.approx(c["pose_checksum"], abs=1e-7)


@pytest.mark.gpu
def test_energy_batch_with_cutoff_matches_the_restatement(gpu, orc, ddd):
    prot = ddd.synth.make_receptor(3000, seed=4)
    off, lig = ddd.synth.make_ligands([40, 21, 33, 1, 64, 2], prot)
    with gpu.Receptor(prot) as rec:
        for rc in (8.0, 12.0, 1000.0):
            eo, ao = orc.energy_batch(prot, off, lig, orc.params(cutoff=rc), with_abs=True)
            e64, a64 = gpu.energy_batch(rec, off, lig, gpu.default_params(precision=gpu.PRECISION_F64, cutoff=rc), with_abs=True)
            e32 = gpu.energy_batch(rec, off, lig, gpu.default_params(cutoff=rc))
            assert np.all(np.abs(e64 - eo) <= 1e-13 * ao + 1e-300) and np.allclose(a64, ao, rtol=1e-12)
            assert np.all(np.abs(e32 - eo) <= 2e-6 * ao + 1e-300)
        # default params: the reference's sum, untouched by the extension
        assert gpu.default_params().cutoff == 0.0
        assert np.array_equal(gpu.energy_batch(rec, off, lig, gpu.default_params(precision=gpu.PRECISION_F64)),
                              gpu.energy_batch(rec, off, lig, gpu.default_params(precision=gpu.PRECISION_F64, cutoff=0.0)))


@pytest.mark.gpu
def test_chains_with_cutoff_replay_the_restatement(gpu, orc, ddd):
    # F64 chains: identical accept trace / draws / retries to the CPU restatement run with the same cutoff
    prot = ddd.synth.make_receptor(2500, seed=6)
    off, lig = ddd.synth.make_ligands([12, 40, 27, 33, 5, 20], prot)
    p64 = gpu.default_params(precision=gpu.PRECISION_F64, cutoff=10.0)
    po = orc.params(cutoff=10.0)
    with gpu.Receptor(prot) as rec:
        poses, stats, trace = gpu.run_chains(rec, off, lig, 2, 120, True, T, p64, seed=8, want_trace=True)
    for c in range(2 * (len(off) - 1)):
        li, r = divmod(c, 2)
        pose_o, st_o, tr_o, _ = orc.chain(prot, lig[off[li]:off[li + 1]], 120, True, T, orc.philox_stream(8, c), po, want_trace=True)
        assert np.array_equal(trace[c], tr_o), c
        assert stats[c]["draws"] == st_o["draws"] and stats[c]["retries"] == st_o["retries"]
        assert np.allclose(gpu.chain_pose(off, 2, poses, li, r), pose_o, rtol=0, atol=1e-9)
        assert stats[c]["final_energy"] == pytest.approx(st_o["final_energy"], rel=1e-10)


@pytest.mark.gpu
def test_cell_list_sums_equal_the_brute_force_sums(gpu, orc, ddd):
    # production precision on a receptor large enough for the cell list to drop most units: the chain's own
    # fp32-summed energy must agree with the fp64 brute-force energy of the same pose (conditioning-aware),
    # whatever the list did; a cutoff that overflows the unit list takes the all-units path and must agree too
    prot = ddd.synth.make_receptor(20000, seed=9)
    off, lig = ddd.synth.make_ligands([40] * 96 + [20, 80, 3], prot)
    with gpu.Receptor(prot) as rec:
        for rc, expect_culling in ((12.0, True), (60.0, False)):
            prm = gpu.default_params(cutoff=rc)
            scr = gpu.Screen(rec, off, lig, 1)
            scr.run(60, True, T, prm, seed=2)
            poses, stats = scr.download()
            units = scr.pair_units()
            scr.run(60, True, T, prm, seed=2)
            poses2, stats2 = scr.download()
            scr.set_segment_steps(16)                                                       # parked and resumed three times: same bits, same counters
            scr.run(60, True, T, prm, seed=2)
            poses3, stats3 = scr.download()
            assert scr.segment_steps() == 16 and np.array_equal(poses, poses3) and np.array_equal(stats, stats3)
            assert np.array_equal(units, scr.pair_units())
            scr.close()
            e64, ab = gpu.energy_batch(rec, off, poses, gpu.default_params(precision=gpu.PRECISION_F64, cutoff=rc), with_abs=True)
            # fp32 rounding scales with the UNSHIFTED terms K|q q|/d (each 1/d is rounded before 1/rc is subtracted)
            _, ab_plain = gpu.energy_batch(rec, off, poses, gpu.default_params(precision=gpu.PRECISION_F64), with_abs=True)
            assert np.array_equal(poses, poses2) and np.array_equal(stats, stats2)           # bit-reproducible
            assert np.all(stats["status"] == 0) and np.all(stats["steps"] == 60)
            assert np.all(np.abs(stats["final_energy"] - e64) <= 1e-13 * ab + 1e-300)
            assert np.all(np.abs(stats["chain_energy"] - e64) <= 1e-6 * ab_plain + 1e-300)
            assert np.all(stats["final_energy"] <= stats["start_energy"])
            full = 61 * 20000                                                               # (start + 60 steps) evaluations x nP
            assert np.all(units > 0)
            if expect_culling:
                assert units.max() < 0.25 * full                                            # most of the receptor is never touched
            else:
                assert units.max() > 0.5 * full                                             # nearly everything is in reach
        # the one-CTA-per-chain engine has no cell list: asking for both is an error, not a silent fallback
        scr = gpu.Screen(rec, off, lig, 1)
        scr.set_chain_threads(128)
        with pytest.raises(gpu.DddError):
            scr.run(5, True, T, gpu.default_params(cutoff=12.0), seed=1)
        scr.close()
