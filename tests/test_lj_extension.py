"""Lennard-Jones extension (params.lj_scale != 0): 12-6 term on top of the reference's Coulomb sum.

The reference has NO Lennard-Jones term and its data model carries no atom types (datatypes.go:18-21), so this
mode has no reference row: parity is against the builder's own CPU restatement only ("parity unpinned"), it is off by
default, and the per-atom parameters travel through extra entry points (ddd_receptor_set_lj, ddd_screen_set_lj).
E = E_coulomb + lj_scale * sum 4 sqrt(eps_p eps_l) (x^6 - x^3),  x = min(((sigma_p + sigma_l)/2)^2 / d^2, lj_cap).
"""
import numpy as np
import pytest

T = 310.15
K = 9e9
SCALE = 2.0e5


def _lj_params(n, seed):
    rng = np.random.default_rng(seed)
    return np.stack([rng.uniform(2.4, 4.0, n), rng.uniform(0.02, 0.3, n)], axis=1)


def _np_energy(prot, lig, plj, llj, scale, cap=4.0, rc=0.0):
    d = np.linalg.norm(prot[:, None, :3] - lig[None, :, :3], axis=-1)
    qq = prot[:, None, 3] * lig[None, :, 3]
    sg = 0.5 * (plj[:, None, 0] + llj[None, :, 0])
    x = np.minimum(sg * sg / np.maximum(d * d, 1e-300), cap)
    t = K * qq / np.maximum(d, 1e-6) + scale * 4 * np.sqrt(plj[:, None, 1]) * np.sqrt(llj[None, :, 1]) * (x ** 6 - x ** 3)
    if rc > 0:
        t = np.where(d < rc, t - K * qq / rc, 0.0)
    return float(t.sum()), float(np.abs(t).sum())


def test_oracle_lj_matches_numpy_and_defaults_off(orc, ddd):
    prot = ddd.synth.make_receptor(700, seed=3)
    off, lig = ddd.synth.make_ligands([19, 30], prot)
    plj = _lj_params(len(prot), 1)
    for i in range(2):
        l = lig[off[i]:off[i + 1]]
        llj = _lj_params(len(l), 10 + i)
        assert orc.params().lj_scale == 0.0
        for rc in (0.0, 9.0):
            e, ab = orc.energy(prot, l, orc.params(plj, llj, lj_scale=SCALE, cutoff=rc), with_abs=True)
            en, abn = _np_energy(prot, l, plj, llj, SCALE, rc=rc)
            assert e == pytest.approx(en, rel=0, abs=1e-11 * abn) and ab == pytest.approx(abn, rel=1e-11)
        assert orc.energy(prot, l, orc.params(plj, llj)) == orc.energy(prot, l)         # lj_scale 0: the reference's sum
        # overlapping atoms: the capped core stays finite
        clash = l.copy(); clash[0, :3] = prot[5, :3]
        assert np.isfinite(orc.energy(prot, clash, orc.params(plj, llj, lj_scale=SCALE)))


@pytest.mark.gpu
def test_lj_energies_and_chains_match_the_restatement(gpu, orc, ddd):
    prot = ddd.synth.make_receptor(2500, seed=6)
    sizes = [12, 40, 27, 33, 5, 20, 1, 64]
    off, lig = ddd.synth.make_ligands(sizes, prot)
    plj, llj = _lj_params(len(prot), 2), _lj_params(len(lig), 3)
    with gpu.Receptor(prot) as rec:
        rec.set_lj(plj[:, 0], plj[:, 1])
        for rc in (0.0, 10.0):
            p64 = gpu.default_params(precision=gpu.PRECISION_F64, lj_scale=SCALE, cutoff=rc)
            p32 = gpu.default_params(lj_scale=SCALE, cutoff=rc)
            scr = gpu.Screen(rec, off, lig, 2)
            scr.set_lj(llj[:, 0], llj[:, 1])
            # energies: a 0-step run returns CalculateEnergy of the start pose (fp64) and the chain's own fp32-summed energy
            scr.run(0, True, T, p32, seed=1)
            _, st0 = scr.download()
            scr.run(100, True, T, p64, seed=8)
            poses, stats = scr.download()
            scr.close()
            for c in range(2 * len(sizes)):
                li, r = divmod(c, 2)
                l, lj = lig[off[li]:off[li + 1]], llj[off[li]:off[li + 1]]
                po = orc.params(plj, lj, lj_scale=SCALE, cutoff=rc)
                eo, ao = orc.energy(prot, l, po, with_abs=True)
                assert abs(st0[c]["start_energy"] - eo) <= 1e-13 * ao + 1e-300
                assert abs(st0[c]["chain_energy"] - eo) <= 3e-6 * ao + 1e-300
                pose_o, st_o = orc.chain(prot, l, 100, True, T, orc.philox_stream(8, c), po)
                for k in ("steps", "accepted", "downhill", "retries", "draws", "status"):
                    assert stats[c][k] == st_o[k], (rc, c, k)
                assert np.allclose(gpu.chain_pose(off, 2, poses, li, r), pose_o, rtol=0, atol=1e-9)
                assert stats[c]["final_energy"] == pytest.approx(st_o["final_energy"], rel=1e-10)
        # LJ changes the answer (the extension is really on) and the default is really off
        scr = gpu.Screen(rec, off, lig, 1)
        scr.set_lj(llj[:, 0], llj[:, 1])
        scr.run(0, True, T, gpu.default_params(precision=gpu.PRECISION_F64), seed=1)
        _, plain = scr.download()
        scr.run(0, True, T, gpu.default_params(precision=gpu.PRECISION_F64, lj_scale=SCALE), seed=1)
        _, with_lj = scr.download()
        scr.close()
        e_ref = gpu.energy_batch(rec, off, lig, gpu.default_params(precision=gpu.PRECISION_F64))
        assert np.all(np.abs(plain["start_energy"] - e_ref) <= 1e-12 * np.abs(e_ref) + 1e-3)
        assert np.any(np.abs(with_lj["start_energy"] - plain["start_energy"]) > 1e-3 * np.abs(plain["start_energy"]))
        # parameters are mandatory, and the one-shot entry points cannot carry them
        scr = gpu.Screen(rec, off, lig, 1)
        with pytest.raises(gpu.DddError):
            scr.run(5, True, T, gpu.default_params(lj_scale=SCALE), seed=1)
        scr.close()
        with pytest.raises(gpu.DddError):
            gpu.energy_batch(rec, off, lig, gpu.default_params(lj_scale=SCALE))
        with pytest.raises(gpu.DddError):
            gpu.run_chains(rec, off, lig, 1, 5, True, T, gpu.default_params(lj_scale=SCALE))


@pytest.mark.gpu
def test_mode_combinations_are_reproducible_and_engine_independent(gpu, ddd):
    # every extension x precision x slot count on a mixed batch with more chains than slots: same seed -> same bits,
    # and the number of resident chains per SM never changes a result
    prot = ddd.synth.make_receptor(2300, seed=3)
    sizes = [3 + (7 * i) % 38 for i in range(40)]
    off, lig = ddd.synth.make_ligands(sizes, prot)
    plj, llj = _lj_params(len(prot), 5), _lj_params(len(lig), 6)
    with gpu.Receptor(prot) as rec:
        rec.set_lj(plj[:, 0], plj[:, 1])
        for kw in ({}, {"precision": gpu.PRECISION_F64}, {"cutoff": 9.0}, {"lj_scale": SCALE},
                   {"cutoff": 9.0, "lj_scale": SCALE, "precision": gpu.PRECISION_F64}):
            ref = None
            for slots in (0, 1, 3, 0):
                scr = gpu.Screen(rec, off, lig, 6)
                scr.set_lj(llj[:, 0], llj[:, 1])
                scr.set_chain_slots(slots)
                scr.run(15, True, T, gpu.default_params(**kw), seed=1)
                p, st = scr.download()
                scr.close()
                assert np.all(st["status"] == 0) and np.all(st["steps"] == 15)
                if ref is None:
                    ref = (p, st)
                else:
                    assert np.array_equal(ref[0], p) and np.array_equal(ref[1], st), (kw, slots)


def test_lj_parameters_from_mol2_types(ddd, sample, sample_types):
    M = ddd.metropolis
    t = sample_types["223l_ligand"].tolist()
    assert t[:6] == ["Cl", "C.3", "C.3", "O.3", "S.3", "C.3"] and len(t) == len(sample["223l_ligand"])
    sg, ep = M.lj_parameters(t)
    assert (sg[0], ep[0]) == M.LJ_BY_ELEMENT["Cl"] and (sg[1], ep[1]) == M.LJ_BY_ELEMENT["C"] and (sg[4], ep[4]) == M.LJ_BY_ELEMENT["S"]
    assert M.lj_parameters(["Xx.9"]) == (np.array([M.LJ_DEFAULT[0]]), np.array([M.LJ_DEFAULT[1]]))
    from pathlib import Path
    f = Path("/root/reference/metropolisMethod/Data/mol2_files/223l_ligand.mol2")
    if f.exists():                                   # build container only: the fixture is what the parser reads
        assert M.ParseMol2Types(f) == t
        assert len(M.ParseMol2Types(f)) == len(M.ParseMol2(f))


@pytest.mark.gpu
def test_lj_on_shipped_data_with_typed_parameters(gpu, orc, ddd, sample, sample_types):
    # 223l protein x shipped ligands, sigma / epsilon from the mol2 atom types: fp64 energies and a chain vs the restatement
    M = ddd.metropolis
    prot = sample["223l_protein"]
    psg, pep = M.lj_parameters(sample_types["223l_protein"].tolist())
    names = ["1a0i", "223l", "421p", "6tol"]
    off, lig = gpu.pack_ligands([sample[n + "_ligand"] for n in names])
    lsg, lep = M.lj_parameters(np.concatenate([sample_types[n + "_ligand"] for n in names]).tolist())
    scale = 1.0e6
    with gpu.Receptor(prot) as rec:
        rec.set_lj(psg, pep)
        scr = gpu.Screen(rec, off, lig, 1)
        scr.set_lj(lsg, lep)
        scr.run(150, True, T, gpu.default_params(precision=gpu.PRECISION_F64, lj_scale=scale), seed=3)
        poses, stats = scr.download()
        scr.close()
    plj = np.stack([psg, pep], axis=1)
    for i, n in enumerate(names):
        l = lig[off[i]:off[i + 1]]
        po = orc.params(plj, np.stack([lsg[off[i]:off[i + 1]], lep[off[i]:off[i + 1]]], axis=1), lj_scale=scale)
        eo, ao = orc.energy(prot, l, po, with_abs=True)
        assert abs(stats[i]["start_energy"] - eo) <= 1e-13 * ao
        assert eo != orc.energy(prot, l)                                        # the 12-6 term is really there
        pose_o, st_o = orc.chain(prot, l, 150, True, T, orc.philox_stream(3, i), po)
        assert stats[i]["accepted"] == st_o["accepted"] and stats[i]["draws"] == st_o["draws"]
        assert np.allclose(gpu.chain_pose(off, 1, poses, i, 0), pose_o, rtol=0, atol=1e-9)
