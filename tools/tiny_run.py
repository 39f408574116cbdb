"""Development: a tiny run that touches every path of the SM engine (for compute-sanitizer memcheck)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ddd_loader
m = ddd_loader.load(); capi, synth = m.capi, m.synth
capi.init([0])
prot = synth.make_receptor(2300, seed=3)
sizes = [3 + (7 * i) % 38 for i in range(40)]
off, lig = synth.make_ligands(sizes, prot)
rng = np.random.default_rng(0)
rec = capi.Receptor(prot); rec.set_lj(rng.uniform(2.4, 4, len(prot)), rng.uniform(0.02, 0.3, len(prot)))
for name, prm, slots in (("f32", capi.default_params(), 0), ("f64", capi.default_params(precision=capi.PRECISION_F64), 2),
                         ("cut", capi.default_params(cutoff=9.0), 3), ("lj", capi.default_params(lj_scale=2e5), 0),
                         ("cut+lj f64", capi.default_params(cutoff=9.0, lj_scale=2e5, precision=capi.PRECISION_F64), 1)):
    scr = capi.Screen(rec, off, lig, 6); scr.set_lj(rng.uniform(2.4, 4, len(lig)), rng.uniform(0.02, 0.3, len(lig)))
    scr.set_chain_slots(slots)
    scr.run(12, True, 310.15, prm, seed=1)
    p, st = scr.download(); r, _ = scr.rmsd(); scr.close()
    print(name, "ok", int(st["accepted"].sum()), int(st["retries"].sum()), float(np.nanmax(r)))
e = capi.energy_batch(rec, off, lig, capi.default_params(cutoff=9.0))
print("energy ok", float(e.sum()))
