"""Development: throughput of the LJ extension on the config-2 shape."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ddd_loader
m = ddd_loader.load(); capi, synth = m.capi, m.synth
capi.init([0])
prot = synth.make_receptor(5000)
off, lig = synth.make_ligands([40] * 1024, prot)
rng = np.random.default_rng(1)
rec = capi.Receptor(prot); rec.set_lj(rng.uniform(2.4, 4.0, 5000), rng.uniform(0.02, 0.3, 5000))
scr = capi.Screen(rec, off, lig, 1); scr.set_lj(rng.uniform(2.4, 4.0, len(lig)), rng.uniform(0.02, 0.3, len(lig)))
for name, prm in (("coulomb", capi.default_params()), ("coulomb+lj", capi.default_params(lj_scale=2e5))):
    scr.run(500, 1, 310.15, prm, seed=1); scr.sync()
    scr.run(2000, 1, 310.15, prm, seed=2); scr.sync()
    ms = scr.last_kernel_ms(); _, st = scr.download()
    tot = float(st["steps"].sum())
    print(f"{name}: {ms:.1f} ms, {tot / ms * 1e3:.3e} MC steps/s, {tot * 5000 * 40 / ms * 1e3:.3e} pairs/s, acc={st['accepted'].mean() / 2000:.4f}")
