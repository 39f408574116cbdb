"""Development check on a multi-GPU box: chains sharded over the devices of ONE process (ddd_init with N ordinals)
reproduce the single-device result bit for bit, F64 and F32, with and without the cutoff extension."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ddd_loader
m = ddd_loader.load(); capi, synth = m.capi, m.synth
import torch
n_dev = torch.cuda.device_count()
prot = synth.make_receptor(3000, seed=5)
sizes = [20 + (11 * i) % 50 for i in range(300)]
off, lig = synth.make_ligands(sizes, prot)
res = {}
for devs in ([0], list(range(n_dev))):
    capi.init(devs)
    rec = capi.Receptor(prot)
    for name, prm in (("f64", capi.default_params(precision=capi.PRECISION_F64)), ("f32", capi.default_params()),
                      ("f32cut", capi.default_params(cutoff=10.0))):
        out, en, picked, st = capi.minimize_batch(rec, off, lig, 240, True, 310.15, 3, prm, seed=4, want_stats=True)
        scr = capi.Screen(rec, off, lig, 2); scr.run(50, True, 310.15, prm, seed=9); p, s = scr.download(); r, _ = scr.rmsd()
        fr = scr.fused_rmsd()
        # the resident pipeline: randomise -> chains -> pick -> RMSD -> argmin, every stage on the device that owns the ligand
        scr.randomize(seed=3); scr.run(40, False, 310.15, prm, seed=11); pp, pe, pk, pr = scr.pick(310.15, seed=11); am = np.array(scr.argmin(1))
        scr.close()
        # the one-shot batches are cut into one block per device
        e = capi.energy_batch(rec, off, lig, prm) if name != "f32cut" else np.zeros(1)
        d, li, pj = capi.closest_atom_batch(rec, off, lig)
        far = lig.copy(); far[:, :3] += 30.0
        sh = capi.shift_closer_batch(rec, off, far, 5.0)
        rz = capi.randomize_pose_batch(off, lig, seed=6)
        rm = capi.rmsd_batch(off, rz, lig, 4)
        res[(len(devs), name)] = (out, en, picked, st, p, s, r, fr, pp, pe, pk, pr, am, e, d, li, pj, sh, rz, rm)
    rec.close(); capi.shutdown()
ok = True
for name in ("f64", "f32", "f32cut"):
    a, b = res[(1, name)], res[(n_dev, name)]
    same = all(np.array_equal(x, y) for x, y in zip(a, b))
    print(f"{name}: 1 device vs {n_dev} devices identical: {same}")
    ok &= same
print("MULTI_GPU_CHECK", "OK" if ok else "FAILED", "devices", n_dev)
sys.exit(0 if ok else 1)
