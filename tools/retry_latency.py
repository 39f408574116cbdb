"""Per-retry latency of the chain kernel: one chain alone on the GPU, with and without a pair of atoms parked just
above MINDISTANCE (which makes most proposals collide)."""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ddd_loader
m = ddd_loader.load(); capi, synth = m.capi, m.synth
capi.init([0])
prot = synth.make_receptor(5000)
off, lig = synth.make_ligands([40], prot)
rec = capi.Receptor(prot)
stuck = lig.copy()
stuck[1, :3] = stuck[0, :3] + np.array([0.505, 0.0, 0.0])
stuck[3, :3] = stuck[2, :3] + np.array([0.0, 0.505, 0.0])
for width in (128, 256, 64):
    for rot in (1, 0):
        res = []
        for name, l in (("clean", lig), ("stuck", stuck)):
            scr = capi.Screen(rec, off, l, 1)
            scr.set_chain_threads(width)
            scr.run(50, rot, 310.15, capi.default_params(), seed=1); scr.sync()
            scr.run(400, rot, 310.15, capi.default_params(), seed=1); scr.sync()
            ms = scr.last_kernel_ms(); _, st = scr.download(); scr.close()
            res.append((ms, int(st["retries"][0]), int(st["steps"][0])))
        (m0, r0, s0), (m1, r1, s1) = res
        R = (m1 - m0) / max(1, r1 - r0) * 1e3
        print(f"width={width} rot={rot}: clean {m0:.2f} ms ({r0} retries, {m0 / s0 * 1e3:.1f} us/step)  stuck {m1:.2f} ms ({r1} retries)  -> {R:.2f} us = {R * 1965:.0f} cycles per retry")
