"""Development: run ONE chain of the config-2 ligand set alone (for ncu source-level profiles of the serial phase).
python tools/one_chain.py <ligand index> <steps> <spec>"""
import sys
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import ddd_loader
m = ddd_loader.load(); capi, synth = m.capi, m.synth
idx, steps, spec = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
capi.init([0])
prot = synth.make_receptor(5000)
off, lig = synth.make_ligands([40] * 1024, prot)
rec = capi.Receptor(prot)
s1 = capi.Screen(rec, np.array([0, 40]), lig[off[idx]:off[idx + 1]], 1, first_chain_id=idx)
if spec[0] == "s": s1.set_chain_slots(int(spec[1:]))
else: s1.set_chain_threads(int(spec[1:]))
s1.run(steps, 1, 310.15, capi.default_params(), seed=2); s1.sync()
_, st = s1.download()
print(f"chain {idx} {spec}: {s1.last_kernel_ms():.2f} ms retries={int(st['retries'][0])} accepted={int(st['accepted'][0])}")
