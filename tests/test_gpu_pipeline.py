"""-m gpu: the screen-resident pipeline -- RandomizeLigandPose -> (ShiftLigandCloserByThreshold) -> chains -> replica pick ->
CalculateRMSD -> SaveMinimumEnergyLigand's argmin without leaving HBM (validateResults.go:14-45, metropolis.go:14-16, :94-111,
plotting.go:109-121), the fused RMSD of the chain epilogue, the best-ever pose tracker (EXTENSION), the receptor cache of the
scalar entry points, stale handles, the chain-status report, and the in-process multi-device path."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, mock_ligand

pytestmark = pytest.mark.gpu
T = 310.15


def _p64(gpu, **kw):
    return gpu.default_params(precision=gpu.PRECISION_F64, **kw)


@pytest.mark.parametrize("rotate", [False, True])
def test_resident_compare_rmsd_pipeline_matches_oracle(gpu, orc, synth_small, rotate):
    # TestMethodRMSD's shape (main.go:116-123: rotate=false) and the rotate=true variant: CompareRMSD per ligand
    prot, off, lig = synth_small
    n, R, width = len(off) - 1, 3, 60
    with gpu.Receptor(prot) as rec:
        scr = gpu.Screen(rec, off, lig, R)
        scr.randomize(seed=5)                                              # validateResults.go:42, resident
        start = scr.download_start()
        assert np.array_equal(start, gpu.randomize_pose_batch(off, lig, seed=5))       # the same kernel, resident or not
        for i in range(n):
            assert np.allclose(start[off[i]:off[i + 1]], orc.randomize_pose(lig[off[i]:off[i + 1]], orc.philox_stream(5, gpu.STREAM_RANDOMIZE_BASE + i)),
                               rtol=0, atol=1e-9)
        scr.run(width, rotate, T, _p64(gpu), seed=5)                       # :43 SimulateEnergyMinimizationParallel: the chains ...
        pose, en, picked, rmsd = scr.pick(T, seed=5)                       # ... and its replica pick, on the device
        idx0, e0 = scr.argmin(gpu.ARGMIN_SAVE_MIN_LIGAND)
        idx1, e1 = scr.argmin(gpu.ARGMIN_RSHINY)
        chain_rmsd = scr.fused_rmsd()
        _, stats = scr.download()
        scr.close()
        e64 = gpu.energy_batch(rec, off, pose, _p64(gpu))
    assert np.allclose(en, e64, rtol=1e-12)                                # energy = CalculateEnergy(picked pose) (:61)
    for i in range(n):
        r_o = orc.compare_rmsd(prot, lig[off[i]:off[i + 1]], R * width, rotate, T, R, 5, i)
        rc, sim_o, st_o, poses_o, pk_o = orc.minimize_parallel(prot, start[off[i]:off[i + 1]], R * width, rotate, T, R, 5, i)
        assert picked[i] == pk_o
        assert np.allclose(pose[off[i]:off[i + 1]], sim_o, rtol=0, atol=1e-9)
        assert rmsd[i] == pytest.approx(r_o, rel=1e-9, abs=1e-9)            # CompareRMSD's return (:44), reference = pose BEFORE randomising
        assert rmsd[i] == pytest.approx(orc.rmsd(pose[off[i]:off[i + 1]], lig[off[i]:off[i + 1]]), rel=1e-13)
        if picked[i] >= 0:
            assert rmsd[i] == chain_rmsd[i * R + picked[i]]                 # the epilogue's value of that chain, bit for bit
    assert (idx0, e0) == orc.min_energy_index(en, 0) and (idx1, e1) == orc.min_energy_index(en, 1)


def test_argmin_rules(gpu, orc, ddd):
    # SaveMinimumEnergyLigand starts from minEnergy 0.0 (plotting.go:111): with only positive energies index 0 "wins" with 0.0;
    # RShinyAppMain starts from 1e20 (main.go:35).  Strict < keeps the FIRST of equal minima.  Two like-charged atoms far
    # apart give positive energies; duplicated ligands give exact ties.
    prot = np.array([[0.0, 0, 0, 1.0], [1.0, 1, 1, 1.0]])
    one = np.array([[6.0, 0, 0, 0.5], [7.0, 0.8, 0, 0.5]])
    neg = one.copy(); neg[:, 3] = -0.5
    for ligs, want_neg in (([one, one * np.array([1.5, 1, 1, 1]), one], False), ([one, neg, one, neg, neg], True)):
        off, x = gpu.pack_ligands(ligs)
        with gpu.Receptor(prot) as rec:
            scr = gpu.Screen(rec, off, x, 2)
            scr.run(0, True, T, _p64(gpu), seed=1)                          # zero steps: every chain returns the start pose
            pose, en, picked, rmsd = scr.pick(T, seed=1)
            got0, got1 = scr.argmin(0), scr.argmin(1)
            scr.close()
        assert np.all(rmsd == 0.0) and np.array_equal(pose, x)
        assert got0 == orc.min_energy_index(en, 0) and got1 == orc.min_energy_index(en, 1)
        if want_neg:
            assert got0[0] == 1 and got1[0] == 1 and en[1] == en[3] == en[4] < 0      # first of the tied minima
        else:
            assert got0 == (0, 0.0) and got1[0] == 1                        # nothing is below 0.0: (0, 0.0) as in Go
    with gpu.Receptor(prot) as rec:
        scr = gpu.Screen(rec, np.array([0, 2]), one, 1)
        scr.run(0, True, T, _p64(gpu), seed=1)
        with pytest.raises(gpu.DddError):
            scr.argmin(0)                                                   # needs the pick first
        scr.close()


def test_resident_shift_then_minimize_is_the_serial_variant(gpu, orc, synth_small):
    # SimulateMultipleLigands (metropolis.go:11-22): ShiftLigandCloserByThreshold for every ligand, then minimise
    prot, off, lig = synth_small
    far = lig.copy(); far[:, :3] += 25.0
    with gpu.Receptor(prot) as rec:
        scr = gpu.Screen(rec, off, far, 2)
        scr.shift_closer(5.0)
        start = scr.download_start()
        assert np.array_equal(start, gpu.shift_closer_batch(rec, off, far, 5.0))
        scr.run(50, True, T, _p64(gpu), seed=3)
        pose, en, picked, rmsd = scr.pick(T, seed=3)
        scr.close()
    rc, xo, eno, sto = orc.simulate_multiple_ligands(prot, off, far, 100, True, T, 2, 3, shift_first=True)
    assert rc == 0 and np.allclose(pose, xo, rtol=0, atol=1e-9) and np.allclose(en, eno, rtol=1e-10)
    for i in range(len(off) - 1):                                            # RMSD reference = the poses as uploaded (before the shift)
        assert rmsd[i] == pytest.approx(orc.rmsd(pose[off[i]:off[i + 1]], far[off[i]:off[i + 1]]), rel=1e-13)


def test_fused_rmsd_equals_the_rmsd_kernel(gpu, orc, ddd):
    prot = ddd.synth.make_receptor(900, seed=3)
    off, lig = ddd.synth.make_ligands([20 + (7 * i) % 61 for i in range(300)], prot)
    with gpu.Receptor(prot) as rec:
        scr = gpu.Screen(rec, off, lig, 2)
        scr.run(30, True, T, gpu.default_params(), seed=2)
        fused = scr.fused_rmsd()
        kern, ms = scr.rmsd()
        poses, _ = scr.download()
        scr.set_chain_threads(64)                                            # one CTA per chain: no fused epilogue
        scr.run(30, True, T, gpu.default_params(), seed=2)
        with pytest.raises(gpu.DddError):
            scr.fused_rmsd()
        kern2, _ = scr.rmsd()
        scr.close()
    assert np.array_equal(fused, kern) and ms > 0
    for c in (0, 1, 77, 599):
        li, r = divmod(c, 2)
        assert fused[c] == pytest.approx(orc.rmsd(gpu.chain_pose(off, 2, poses, li, r), lig[off[li]:off[li + 1]]), rel=1e-13)
    assert np.all(np.isfinite(kern2))


@pytest.mark.parametrize("temperature", [310.15, 3.0e8])
def test_best_ever_tracker_matches_its_oracle(gpu, orc, synth_small, temperature):
    # EXTENSION (no reference row): at the reference temperature every accepted move is downhill and best == final; with
    # uphill accepts (T = 3e8) the best pose is an earlier one and has to be saved when it is left
    prot, off, lig = synth_small
    n, R = len(off) - 1, 2
    with gpu.Receptor(prot) as rec:
        scr = gpu.Screen(rec, off, lig, R)
        scr.set_track_best(True)
        scr.run(150, True, temperature, _p64(gpu), seed=9)
        poses, stats = scr.download()
        best, best_e = scr.download_best()
        scr.set_chain_threads(64)
        with pytest.raises(gpu.DddError):                                    # the tracker lives in the SM-persistent engine
            scr.run(10, True, temperature, _p64(gpu), seed=9)
        scr.close()
    differs = 0
    for c in range(n * R):
        li, r = divmod(c, R)
        pose_o, st_o, best_o, be_o = orc.chain_best(prot, lig[off[li]:off[li + 1]], 150, True, temperature, orc.philox_stream(9, c))
        assert stats[c]["accepted"] == st_o["accepted"] and stats[c]["draws"] == st_o["draws"]
        assert best_e[c] == pytest.approx(be_o, rel=1e-10)
        assert np.allclose(gpu.chain_pose(off, R, best, li, r), best_o, rtol=0, atol=1e-9)
        assert best_e[c] <= stats[c]["final_energy"] + 1e-9 * abs(stats[c]["final_energy"])
        differs += not np.array_equal(gpu.chain_pose(off, R, best, li, r), gpu.chain_pose(off, R, poses, li, r))
    if temperature < 1e3:
        assert differs == 0 and np.allclose(best_e, stats["final_energy"], rtol=1e-10)
    else:
        assert differs > 0


def test_best_ever_tracker_production_precision(gpu, ddd):
    prot = ddd.synth.make_receptor(1500, seed=2)
    off, lig = ddd.synth.make_ligands([24] * 64, prot)
    with gpu.Receptor(prot) as rec:
        scr = gpu.Screen(rec, off, lig, 1)
        scr.run(200, True, T, gpu.default_params(), seed=3)
        p0, s0 = scr.download()
        scr.set_track_best(True)
        scr.run(200, True, T, gpu.default_params(), seed=3)
        p1, s1 = scr.download()
        best, best_e = scr.download_best()
        scr.close()
    assert np.array_equal(p0, p1) and np.array_equal(s0, s1)                 # tracking changes nothing else
    assert np.array_equal(best, p1) and np.array_equal(best_e, s1["chain_energy"])   # greedy regime: best == final


def test_large_ligand_uses_the_staged_uniforms(gpu, orc, ddd):
    # Rshiny/MCdata ligands reach 140 atoms: the proposal's staged jitter uniforms cover ligands up to 512 atoms
    prot = ddd.synth.make_receptor(400, seed=6)
    off, lig = ddd.synth.make_ligands([140, 97, 85], prot)
    with gpu.Receptor(prot) as rec:
        poses, stats, trace = gpu.run_chains(rec, off, lig, 1, 25, True, T, _p64(gpu), seed=4, want_trace=True)
    for c in range(3):
        pose_o, st_o, tr_o, _ = orc.chain(prot, lig[off[c]:off[c + 1]], 25, True, T, orc.philox_stream(4, c), want_trace=True)
        assert np.array_equal(trace[c], tr_o) and stats[c]["draws"] == st_o["draws"] and stats[c]["retries"] == st_o["retries"]
        assert np.allclose(gpu.chain_pose(off, 1, poses, c, 0), pose_o, rtol=0, atol=1e-9)


def test_clamp_free_margin_follows_clamp_distance(gpu, orc, ddd):
    # a caller-set clamp_distance far above the default 1e-6: groups whose box is within the clamp distance of the pose must
    # keep the clamp (the chain's own fp32-summed start energy = the oracle's clamped sum)
    prot = ddd.synth.make_receptor(700, seed=9)
    rng = np.random.default_rng(3)
    ligs = []
    for i in range(12):                                                      # ligand atoms sit 0.05 .. 0.6 A from receptor atoms
        k = rng.choice(len(prot), size=10, replace=False)
        x = prot[k, :3] + rng.normal(size=(10, 3)) * 0.2
        ligs.append(np.concatenate([x, rng.normal(0, 0.3, size=(10, 1))], axis=1))
    off, lig = gpu.pack_ligands(ligs)
    for clamp in (0.3, 0.9):
        po = orc.params(clamp_distance=clamp)
        with gpu.Receptor(prot) as rec:
            scr = gpu.Screen(rec, off, lig, 1)
            scr.run(0, True, T, gpu.default_params(clamp_distance=clamp), seed=1)
            _, st = scr.download()
            scr.close()
        for i in range(12):
            e, ab = orc.energy(prot, ligs[i], po, with_abs=True)
            assert abs(st[i]["start_energy"] - e) <= 1e-13 * ab
            assert abs(st[i]["chain_energy"] - e) <= 2e-6 * ab, (clamp, i)


def test_receptor_cache(gpu, ddd):
    lib = gpu.lib()
    prot = ddd.synth.make_receptor(300, seed=1)
    a = gpu.Receptor(prot, cached=True)
    b = gpu.Receptor(prot.copy(), cached=True)                               # same content, other buffer: same resident receptor
    assert a.handle.value == b.handle.value
    other = prot.copy(); other[7, 0] += 1e-9
    c = gpu.Receptor(other, cached=True)
    assert c.handle.value != a.handle.value
    off, lig = ddd.synth.make_ligands([9, 14], prot)
    assert np.array_equal(gpu.energy_batch(a, off, lig), gpu.energy_batch(gpu.Receptor(prot), off, lig))
    assert lib.ddd_receptor_destroy(a.handle) != 0                           # cache-owned: destroy is refused
    ha = a.handle.value
    for r in (a, b, c):
        r.close()
    again = gpu.Receptor(prot, cached=True)                                  # still resident after every user released it
    assert again.handle.value == ha
    again.close()
    held = [gpu.Receptor(ddd.synth.make_receptor(40 + i, seed=i), cached=True) for i in range(12)]     # more than the cache keeps
    for i, r in enumerate(held):                                             # entries in use are never evicted
        assert lib.ddd_receptor_num_atoms(r.handle) == 40 + i
    for r in held:
        r.close()


def test_chain_status_is_reported(gpu):
    prot = np.array([[5.0, 5, 5, 1.0], [-3.0, 2, 1, -0.5]])
    bad = np.array([[0.0, 0, 0, 0.1], [0.01, 0, 0, -0.1]])                   # can never get 0.5 A apart: the reference would spin (:155)
    ok = mock_ligand()
    off, lig = gpu.pack_ligands([ok, bad, ok])
    with gpu.Receptor(prot) as rec:
        gpu.minimize_batch(rec, off, lig, 40, False, T, 2, _p64(gpu, max_retries=5), seed=1)
        bits, n = gpu.last_chain_status()
        assert bits & gpu.STATUS_RETRY_LIMIT and n == 2                      # both replicas of the impossible ligand
        gpu.minimize_batch(rec, off[:2], lig[:2], 40, False, T, 2, _p64(gpu), seed=1)
        assert gpu.last_chain_status() == (0, 0)


_STALE = r"""
import sys, numpy as np
sys.path.insert(0, %r)
import ddd_loader
m = ddd_loader.load(); capi, synth = m.capi, m.synth
capi.init([0])
prot = synth.make_receptor(200, seed=1)
off, lig = synth.make_ligands([8, 11], prot)
rec = capi.Receptor(prot)
scr = capi.Screen(rec, off, lig, 1)
e0 = capi.energy_batch(rec, off, lig)
assert capi.init([0]) == 1                       # same device list: a no-op, handles stay valid
assert np.array_equal(capi.energy_batch(rec, off, lig), e0)
capi.shutdown()
capi.init([0])                                   # a new initialisation: the old handles are stale
for f in (lambda: capi.energy_batch(rec, off, lig), lambda: scr.run(5, True, 310.15), lambda: capi.Screen(rec, off, lig, 1)):
    try:
        f(); raise SystemExit("stale handle accepted")
    except capi.DddError as e:
        assert "stale" in str(e), e
scr.close(); rec.close()                         # freeing stale handles is fine
rec = capi.Receptor(prot)
assert np.array_equal(capi.energy_batch(rec, off, lig), e0)
rec.close(); capi.shutdown()
print("STALE_OK")
"""


def test_stale_handles_after_shutdown():
    res = subprocess.run([sys.executable, "-c", _STALE % str(ROOT)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0 and "STALE_OK" in res.stdout, res.stdout + res.stderr


def test_in_process_multi_device_matches_one_device():
    # the path main.go would use: ddd_init with every device, chains sharded inside ONE process (no collective)
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run tools/multi_gpu_check.py under gpurun --gpus 2; log in profiles/)")
    env = dict(os.environ)
    res = subprocess.run([sys.executable, str(ROOT / "tools" / "multi_gpu_check.py")], capture_output=True, text=True, timeout=1200, env=env)
    assert res.returncode == 0 and "MULTI_GPU_CHECK OK" in res.stdout, res.stdout + res.stderr


def test_concurrent_callers_get_the_serial_results(gpu, ddd):
    # goroutines call the batch entry points from many OS threads (metropolis.go:43, :100): the library lock is released while
    # a caller blocks on the device, so callers overlap their staging with each other's kernels -- and results must not change
    import threading
    prot = ddd.synth.make_receptor(1500, seed=2)
    off, lig = ddd.synth.make_ligands([24] * 96, prot)
    blocks = [(0, 32), (32, 64), (64, 96)]
    with gpu.Receptor(prot) as rec:
        def call(b):
            lo, hi = b
            o = off[lo:hi + 1] - off[lo]
            return gpu.minimize_batch(rec, o, lig[off[lo]:off[hi]], 600, True, T, 3, gpu.default_params(), seed=7, first_ligand_index=lo)
        serial = [call(b) for b in blocks]
        results = [None] * len(blocks)
        def worker(k):
            results[k] = call(blocks[k])
        for _ in range(3):
            threads = [threading.Thread(target=worker, args=(k,)) for k in range(len(blocks))]
            for t in threads:
                t.start()
            for t in threads:
                t.join()
            for a, b in zip(serial, results):
                assert all(np.array_equal(x, y) for x, y in zip(a, b))
        whole = gpu.minimize_batch(rec, off, lig, 600, True, T, 3, gpu.default_params(), seed=7)
        assert np.array_equal(np.concatenate([r[0] for r in serial]), whole[0])
