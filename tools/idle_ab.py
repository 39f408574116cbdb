"""Development: one process, the timings that idle-poll changes move: a lone chain with/without retries, config 2 (seeds), a config 3 shard, config 4 cutoff."""
import os, sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ddd_loader
m = ddd_loader.load(); capi, synth = m.capi, m.synth
capi.init([0])
tag = os.environ.get("DDD_MC_LIB", "default")
prm = capi.default_params()
prot = synth.make_receptor(5000)
off, lig = synth.make_ligands([40] * 1024, prot)
rec = capi.Receptor(prot)
out = []
for idx in (945, 96):
    s1 = capi.Screen(rec, np.array([0, 40]), lig[off[idx]:off[idx + 1]], 1, first_chain_id=idx)
    s1.run(2000, 1, 310.15, prm, seed=2); s1.sync()
    s1.run(10000, 1, 310.15, prm, seed=2); s1.sync()
    out.append(f"lone{idx} {s1.last_kernel_ms():.1f}")
    s1.close()
scr = capi.Screen(rec, off, lig, 1)
scr.run(1000, 1, 310.15, prm, seed=1); scr.sync()
for seed in (2, 200, 203, 401):
    scr.run(10000, 1, 310.15, prm, seed=seed); scr.sync()
    out.append(f"c2s{seed} {scr.last_kernel_ms():.1f}")
scr.close()
off3, lig3 = synth.make_ligands([40] * 512, prot)
scr = capi.Screen(rec, off3, lig3, 8)
scr.run(500, 1, 310.15, prm, seed=1); scr.sync()
scr.run(10000, 1, 310.15, prm, seed=200); scr.sync()
out.append(f"c3shard {scr.last_kernel_ms():.1f}")
scr.close(); rec.close()
prot4 = synth.make_receptor(50000)
off4, lig4 = synth.make_ligands([40] * 1024, prot4, workers=8)
rec4 = capi.Receptor(prot4)
scr = capi.Screen(rec4, off4, lig4, 1)
p4 = capi.default_params(cutoff=12.0)
scr.run(1000, 1, 310.15, p4, seed=1); scr.sync()
scr.run(10000, 1, 310.15, p4, seed=200); scr.sync()
out.append(f"c4cut {scr.last_kernel_ms():.1f}")
print(tag.split('/')[-1].ljust(16), "  ".join(out), flush=True)
