import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ddd_loader
m = ddd_loader.load(); gpu, synth = m.capi, m.synth
gpu.init([0])
prot = synth.make_receptor(20000, seed=9)
off, lig = synth.make_ligands([40] * 96 + [20, 80, 3], prot)
rec = gpu.Receptor(prot)
for rc in (12.0, 60.0):
    prm = gpu.default_params(cutoff=rc)
    for steps in (0, 60):
        scr = gpu.Screen(rec, off, lig, 1)
        scr.run(steps, True, 310.15, prm, seed=2)
        poses, stats = scr.download(); units = scr.pair_units(); scr.close()
        e64, ab = gpu.energy_batch(rec, off, poses, gpu.default_params(precision=gpu.PRECISION_F64, cutoff=rc), with_abs=True)
        e32 = gpu.energy_batch(rec, off, poses, gpu.default_params(cutoff=rc))
        r = np.abs(stats["chain_energy"] - e64) / np.maximum(ab, 1e-300)
        r2 = np.abs(e32 - e64) / np.maximum(ab, 1e-300)
        bad = np.argsort(r)[-4:]
        print(f"rc={rc} steps={steps}: max chain ratio {r.max():.3e} at {bad} nl={np.diff(off)[bad]} ratios={r[bad]}; energy_batch f32 max ratio {r2.max():.3e}; units/eval min/max {units.min() / max(1, steps + 1):.0f} {units.max() / max(1, steps + 1):.0f}")
